"""Prompt envelope around the spec — decides the token stream the engine prefills.

Slot order is the reference's (skills/adversarial-spec/scripts/prompts.py:233-241):
system prompt -> "This is round r ..." -> spec -> context -> focus -> instruction,
and the protocol markers ([AGREE], [SPEC]...[/SPEC]) are the reference's, because
reply parsing keys on them (models.py:149-160).  The wording below is this
repo's own stand-in: when the reference's scripts directory is importable (set
ADVSPEC_REFERENCE_SCRIPTS, or drop this package beside it — INTEGRATION.md) its
``prompts`` module is used instead so the prompt bytes are exactly the reference's.
"""

from __future__ import annotations

import importlib.util
import os
from pathlib import Path
from typing import Optional, Sequence

SYSTEM_PROMPT_PRD = """You are a senior product manager taking part in adversarial spec development: several reviewers challenge a Product Requirements Document until it is ready to hand to engineering.

Read the document as the person who will be held to it. Look for: a problem statement a new reader can follow; named users and the jobs they are trying to do; user stories with acceptance criteria that can be tested; success metrics with numbers and a way to measure them; explicit scope and non-goals; dependencies, risks and open questions with an owner each; a rollout and measurement plan.

How to answer:
1. If anything material is missing, vague or contradictory, write a critique that names each problem, why it matters, and what to change. Then output the complete revised document between [SPEC] and [/SPEC] tags.
2. If the document is complete and you would sign off on it, output exactly [AGREE] on its own line, then the final document between [SPEC] and [/SPEC] tags.

Be demanding. Agreement is earned by the document, not by the number of rounds."""

SYSTEM_PROMPT_TECH = """You are a senior software architect taking part in adversarial spec development: several reviewers challenge a Technical Specification until it can be built and operated without guesswork.

Read the document as the engineer who will be paged for it. Look for: a clear statement of what is being built and why; architecture and component responsibilities; data model, storage and migration plan; API contracts with request, response and error shapes; consistency, concurrency and failure handling; security boundaries, authentication and authorisation; performance targets with numbers (latency, throughput, availability) and how they are met; observability, deployment, rollback and capacity; testing strategy; alternatives considered and open questions.

How to answer:
1. If anything material is missing, ambiguous or wrong, write a critique that names each problem, why it matters, and what to change. Then output the complete revised document between [SPEC] and [/SPEC] tags.
2. If the specification is complete and production-ready, output exactly [AGREE] on its own line, then the final document between [SPEC] and [/SPEC] tags.

Be rigorous. Do not agree while a competent engineer would still have to guess."""

SYSTEM_PROMPT_GENERIC = """You are an experienced reviewer taking part in adversarial spec development: several reviewers challenge a specification until it is complete, consistent and actionable.

Look for missing requirements, ambiguity, contradictions, untestable statements, unstated assumptions and unowned risks.

How to answer:
1. If you find material problems, critique them specifically, then output the complete revised document between [SPEC] and [/SPEC] tags.
2. If the document is ready, output exactly [AGREE] on its own line, then the final document between [SPEC] and [/SPEC] tags.

Be demanding. Do not agree out of politeness."""

REVIEW_PROMPT_TEMPLATE = """This is round {round} of adversarial spec development.

Here is the current {doc_type_name}:

{spec}

{context_section}
{focus_section}
Review this document against your criteria. Either critique and revise it, or answer [AGREE] if it is production-ready."""

PRESS_PROMPT_TEMPLATE = """This is round {round} of adversarial spec development. You previously agreed with this document.

Here is the current {doc_type_name}:

{spec}

{context_section}
**Confirm your agreement by reviewing the ENTIRE document again.**

Before answering [AGREE] you must: say that you have read every section; list at least three sections you checked and what you verified in each; explain why the document is complete; and name any remaining concern, however small.

If this second pass turns up a problem, give your critique instead. If you still agree, output your verification, then [AGREE] on its own line, then the final document between [SPEC] and [/SPEC] tags."""

EXPORT_TASKS_PROMPT = """Read this {doc_type_name} and turn it into a list of actionable tasks.

Document:
{spec}

Write every task in exactly this format:
[TASK]
title: <short task title>
type: <user-story | bug | task | spike>
priority: <high | medium | low>
description: <what has to be done>
acceptance_criteria:
- <criterion 1>
- <criterion 2>
[/TASK]

Cover user stories, technical requirements, risks (as spikes) and non-functional requirements. Every
actionable statement in the document should end up in some task."""

PRESERVE_INTENT_PROMPT = """**PRESERVE THE AUTHOR'S INTENT**
Treat every existing requirement as deliberate. Do not delete or substantially rewrite content unless you quote it, state the concrete harm it causes, and show that your change fixes that harm. Prefer adding or clarifying over removing. List every removal separately with its justification."""

FOCUS_AREAS = {
    "security": "**CRITICAL FOCUS: SECURITY**\nPut security first: authentication and authorisation, input validation, secrets, data protection at rest and in transit, abuse cases, audit trails and the blast radius of a compromise.",
    "scalability": "**CRITICAL FOCUS: SCALABILITY**\nPut scale first: growth assumptions, bottlenecks, partitioning, statelessness, back-pressure, hot keys, capacity planning and cost at 10x and 100x load.",
    "performance": "**CRITICAL FOCUS: PERFORMANCE**\nPut performance first: latency and throughput targets, critical-path analysis, caching, batching, data locality, tail latency and how each target will be measured.",
    "ux": "**CRITICAL FOCUS: USER EXPERIENCE**\nPut the user first: flows, error states, empty states, accessibility, consistency, feedback, recovery from mistakes and time to first value.",
    "reliability": "**CRITICAL FOCUS: RELIABILITY**\nPut reliability first: failure modes, retries and idempotency, timeouts, degradation, data durability, recovery objectives, on-call signals and runbooks.",
    "cost": "**CRITICAL FOCUS: COST**\nPut cost first: infrastructure and third-party spend, unit economics, cost drivers that grow with usage, and cheaper designs that meet the same requirements.",
}

PERSONAS = {
    "security-engineer": "You are a security engineer reviewing this document in adversarial spec development. Think like an attacker: find every trust boundary, missing control and data exposure, and say how each would be exploited and fixed.",
    "oncall-engineer": "You are the on-call engineer who will carry the pager for this system, reviewing it in adversarial spec development. Ask what breaks at 3 a.m., how you would know, and how you would fix it without the author.",
    "junior-developer": "You are a junior developer who must implement this document, reviewing it in adversarial spec development. Flag everything you would have to ask someone about: undefined terms, missing steps and unstated assumptions.",
    "qa-engineer": "You are a QA engineer reviewing this document in adversarial spec development. Demand testable statements: edge cases, acceptance criteria, test data, and the behaviour at every boundary.",
    "site-reliability": "You are a site reliability engineer reviewing this document in adversarial spec development. Press on objectives, capacity, rollout, rollback, observability and failure isolation.",
    "product-manager": "You are a product manager reviewing this document in adversarial spec development. Press on user value, success metrics, scope control and the order in which things ship.",
    "data-engineer": "You are a data engineer reviewing this document in adversarial spec development. Press on schemas, lineage, quality checks, retention, backfills and downstream consumers.",
    "mobile-developer": "You are a mobile developer reviewing this document in adversarial spec development. Press on offline behaviour, payload size, versioning, battery and flaky networks.",
    "accessibility-specialist": "You are an accessibility specialist reviewing this document in adversarial spec development. Press on assistive technology, contrast, focus order, motion and inclusive language.",
    "legal-compliance": "You are a legal and compliance reviewer in adversarial spec development. Press on personal data, consent, retention, jurisdiction, contractual commitments and audit evidence.",
}


def _own_get_system_prompt(doc_type: str, persona: Optional[str] = None) -> str:
    if persona:
        key = persona.lower().replace(" ", "-").replace("_", "-")
        if key in PERSONAS:
            return PERSONAS[key]
        return (f"You are a {persona} participating in adversarial spec development. Review the document "
                f"from your professional perspective and critique any issues you find.")
    if doc_type == "prd":
        return SYSTEM_PROMPT_PRD
    if doc_type == "tech":
        return SYSTEM_PROMPT_TECH
    return SYSTEM_PROMPT_GENERIC


def _own_get_doc_type_name(doc_type: str) -> str:
    return {"prd": "Product Requirements Document", "tech": "Technical Specification"}.get(doc_type, "specification")


def _load_reference_prompts():
    root = os.environ.get("ADVSPEC_REFERENCE_SCRIPTS")
    if not root:
        return None
    path = Path(root) / "prompts.py"
    if not path.exists():
        return None
    spec = importlib.util.spec_from_file_location("_advspec_reference_prompts", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_ref = _load_reference_prompts()
SOURCE = "reference" if _ref is not None else "stand-in"
if _ref is not None:
    FOCUS_AREAS = _ref.FOCUS_AREAS
    PERSONAS = _ref.PERSONAS
    PRESERVE_INTENT_PROMPT = _ref.PRESERVE_INTENT_PROMPT
    PRESS_PROMPT_TEMPLATE = _ref.PRESS_PROMPT_TEMPLATE
    EXPORT_TASKS_PROMPT = _ref.EXPORT_TASKS_PROMPT
    REVIEW_PROMPT_TEMPLATE = _ref.REVIEW_PROMPT_TEMPLATE
    get_system_prompt = _ref.get_system_prompt
    get_doc_type_name = _ref.get_doc_type_name
else:
    get_system_prompt = _own_get_system_prompt
    get_doc_type_name = _own_get_doc_type_name


def build_messages(spec: str, round_num: int, doc_type: str, press: bool = False, focus: Optional[str] = None,
                   persona: Optional[str] = None, context: Optional[str] = None,
                   preserve_intent: bool = False) -> tuple[str, str]:
    """(system_prompt, user_message) exactly as call_single_model assembles them
    (skills/adversarial-spec/scripts/models.py:482-503)."""
    system_prompt = get_system_prompt(doc_type, persona)
    focus_section = ""
    if focus and focus.lower() in FOCUS_AREAS:
        focus_section = FOCUS_AREAS[focus.lower()]
    elif focus:
        focus_section = (f"**CRITICAL FOCUS: {focus.upper()}**\n"
                         f"Prioritize analysis of {focus} concerns above all else.")
    if preserve_intent:
        focus_section = PRESERVE_INTENT_PROMPT + "\n\n" + focus_section
    template = PRESS_PROMPT_TEMPLATE if press else REVIEW_PROMPT_TEMPLATE
    user_message = template.format(round=round_num, doc_type_name=get_doc_type_name(doc_type), spec=spec,
                                   focus_section=focus_section, context_section=context if context else "")
    return system_prompt, user_message


PERSONA_TAIL_HEADER = "For this review, answer from the following perspective.\n"


def persona_tail(doc_type: str, persona: str) -> str:
    """What a per-opponent persona adds at the END of the user message: the persona's own prompt text
    (`get_system_prompt(doc_type, persona)`, prompts.py:290-304) under a one-line header."""
    return "\n\n" + PERSONA_TAIL_HEADER + get_system_prompt(doc_type, persona)


def build_panel_messages(spec: str, round_num: int, doc_type: str, press: bool, focus: Optional[str],
                         personas: Sequence[Optional[str]], context: Optional[str] = None,
                         preserve_intent: bool = False) -> tuple[str, list[str]]:
    """(system_prompt, one user message per opponent) for a panel whose opponents may each have a persona of
    their own (SURVEY.md §8(f4) — not in the reference, where `--persona` is one flag for the whole panel,
    debate.py:835, and replaces the SYSTEM prompt, i.e. the first tokens of the prompt).

    One persona for everyone (or none): exactly `build_messages`, the reference's layout.  Different personas:
    the variable text moves BEHIND the document — the system prompt is the doc type's default for every
    opponent and each persona's text closes that opponent's user message — so the prompts share every token up
    to the end of the review instruction and the engine prefills that prefix once for the whole panel."""
    personas = list(personas)
    if len(set(personas)) <= 1:
        system_prompt, user = build_messages(spec, round_num, doc_type, press, focus,
                                             personas[0] if personas else None, context, preserve_intent)
        return system_prompt, [user] * len(personas)
    system_prompt, user = build_messages(spec, round_num, doc_type, press, focus, None, context, preserve_intent)
    return system_prompt, [user + (persona_tail(doc_type, p) if p else "") for p in personas]
