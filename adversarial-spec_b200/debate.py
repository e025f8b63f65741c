#!/usr/bin/env python3
"""`debate.py critique` / `debate.py providers` over the local B200 engine.

Drop-in for the two commands of the reference CLI that sit on the hot path
(skills/adversarial-spec/scripts/debate.py: create_parser :358-432, main
:1025-1055, run_critique :798-888, output_results :891-973): same flags, same
stdout JSON/text, same stderr progress lines, same exit codes (empty stdin 1,
no models / missing key 2, model errors 0).  The other nine actions are config
CRUD or single remote calls with no fan-out and answer with a
pointer to the reference CLI.
"""

from __future__ import annotations

import argparse
import os
import json
import sys
from datetime import datetime
from pathlib import Path
from typing import Any, Optional

if __package__ in (None, ""):  # executed as a script: make the hyphenated package importable
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    import advspec_loader

    advspec_loader.load()
    from advspec_b200 import models as _models, providers as _providers, session as _session
    from advspec_b200.envelope import EXPORT_TASKS_PROMPT, get_doc_type_name
else:
    from . import models as _models, providers as _providers, session as _session
    from .envelope import EXPORT_TASKS_PROMPT, get_doc_type_name

call_models_parallel = _models.call_models_parallel  # tests patch ``debate.call_models_parallel``
completion = _models.completion  # second call site of the reference (debate.py:715, export-tasks)
extract_tasks = _models.extract_tasks
is_o_series_model = _models.is_o_series_model
cost_tracker = _models.cost_tracker
load_context_files = _models.load_context_files
SessionState = _session.SessionState
save_checkpoint = _session.save_checkpoint
DEFAULT_CODEX_REASONING = _providers.DEFAULT_CODEX_REASONING

ACTIONS = ["critique", "providers", "send-final", "diff", "export-tasks", "focus-areas", "personas",
           "profiles", "save-profile", "sessions", "bedrock"]


def create_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Adversarial spec debate with multiple LLMs",
                                formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("action", choices=ACTIONS, help="Action to perform")
    p.add_argument("profile_name", nargs="?", help="Profile name (for save-profile action) or bedrock subcommand")
    p.add_argument("--models", "-m", default=None,
                   help="Comma-separated list of models (e.g., b200/llama-3-8b,b200/llama-3-8b,gpt-4o)")
    p.add_argument("--doc-type", "-d", choices=["prd", "tech"], default="tech",
                   help="Document type: prd or tech (default: tech)")
    p.add_argument("--round", "-r", type=int, default=1, help="Current round number")
    p.add_argument("--rounds", type=int, default=1, help="Total rounds completed (used with send-final)")
    p.add_argument("--json", "-j", action="store_true", help="Output as JSON")
    p.add_argument("--show-cost", action="store_true", help="Show cost summary after critique")
    p.add_argument("--telegram", "-t", action="store_true", help="Send Telegram notifications (not in this build)")
    p.add_argument("--poll-timeout", type=int, default=60, help="Seconds to wait for Telegram reply (default: 60)")
    p.add_argument("--press", "-p", action="store_true",
                   help="Press models to confirm they read the full document (anti-laziness check)")
    p.add_argument("--focus", "-f", help="Focus area for critique (security, scalability, performance, ux, reliability, cost)")
    p.add_argument("--persona", help="Persona for critique (security-engineer, oncall-engineer, junior-developer, etc.)")
    p.add_argument("--context", "-c", action="append", default=[],
                   help="Additional context file(s) to include (can be used multiple times)")
    p.add_argument("--preserve-intent", action="store_true",
                   help="Require explicit justification for any removal or substantial modification")
    p.add_argument("--session", "-s", help="Session ID for state persistence (enables checkpointing and resume)")
    p.add_argument("--resume", help="Resume a previous session by ID")
    p.add_argument("--profile", help="Load settings from a saved profile (not in this build)")
    p.add_argument("--previous", help="Previous spec file (for diff action)")
    p.add_argument("--current", help="Current spec file (for diff action)")
    p.add_argument("--codex-reasoning", default=DEFAULT_CODEX_REASONING, choices=["low", "medium", "high", "xhigh"],
                   help=f"Reasoning effort for Codex CLI models (default: {DEFAULT_CODEX_REASONING})")
    p.add_argument("--codex-search", action="store_true", help="Enable web search for Codex CLI models")
    p.add_argument("--region", help="AWS region for Bedrock (e.g., us-east-1)")
    p.add_argument("bedrock_arg", nargs="?", help="Additional argument for bedrock subcommands")
    p.add_argument("--timeout", type=int, default=600,
                   help="Timeout in seconds for model API/CLI calls (default: 600 = 10 minutes)")
    return p


def parse_models(args: argparse.Namespace) -> list[str]:
    if args.models is None:
        default_model = _providers.get_default_model()
        if default_model is None:
            print("Error: No API keys configured and no models specified.", file=sys.stderr)
            print("\nLocal models need no key: --models b200/llama-3-8b,b200/llama-3-8b", file=sys.stderr)
            print("\nRun 'python3 debate.py providers' to see which keys are set.", file=sys.stderr)
            sys.exit(2)
        args.models = default_model
    models = [m.strip() for m in args.models.split(",") if m.strip()]
    if not models:
        print("Error: No models specified", file=sys.stderr)
        sys.exit(1)
    return models


def validate_models_before_run(models: list[str], bedrock_mode: bool) -> None:
    if bedrock_mode:
        return
    _, invalid = _providers.validate_model_credentials(models)
    if invalid:
        print("Error: The following models lack required API keys:", file=sys.stderr)
        for m in invalid:
            print(f"  - {m} ({_providers.required_key_hint(m)})", file=sys.stderr)
        print("\nRun 'python3 debate.py providers' to see which API keys are configured.", file=sys.stderr)
        sys.exit(2)


def read_spec_from_stdin() -> str:
    """stdin, stripped (reference debate.py:778).  Under torchrun with ADVSPEC_TP=k the k ranks inherit ONE
    stdin, so only rank 0 reads it and the text is broadcast to the other ranks of the tensor-parallel group
    (every rank must run the identical round)."""
    try:
        tp = int(os.environ.get("ADVSPEC_TP", "1"))
    except ValueError:
        tp = 1
    if tp <= 1:
        return sys.stdin.read().strip()
    rank = int(os.environ.get("RANK", "0"))
    text = sys.stdin.read().strip() if rank == 0 else None
    return _models.runtime.broadcast_object(text, src=0)


def load_or_resume_session(args: argparse.Namespace, models: list[str]):
    session_state = None
    if args.resume:
        try:
            session_state = SessionState.load(args.resume)
        except FileNotFoundError as e:
            print(f"Error: {e}", file=sys.stderr)
            sys.exit(2)
        print(f"Resuming session '{args.resume}' at round {session_state.round}", file=sys.stderr)
        spec = session_state.spec
        args.round, args.doc_type = session_state.round, session_state.doc_type
        args.models = ",".join(session_state.models)
        args.focus = session_state.focus or args.focus
        args.persona = session_state.persona or args.persona
        args.preserve_intent = session_state.preserve_intent or args.preserve_intent
        models = session_state.models
    else:
        spec = read_spec_from_stdin()
        if not spec:
            print("Error: No spec provided via stdin", file=sys.stderr)
            sys.exit(1)
    if args.session and not session_state:
        session_state = SessionState(session_id=args.session, spec=spec, round=args.round, doc_type=args.doc_type,
                                     models=models, focus=args.focus, persona=args.persona,
                                     preserve_intent=args.preserve_intent, created_at=datetime.now().isoformat())
        session_state.save()
        print(f"Session '{args.session}' created", file=sys.stderr)
    return spec, session_state, models


def run_critique(args, spec: str, models: list[str], session_state, context: Optional[str],
                 bedrock_mode: bool, bedrock_region: Optional[str]) -> None:
    mode = "pressing for confirmation" if args.press else "critiquing"
    extras = "".join([f" (focus: {args.focus})" if args.focus else "",
                      f" (persona: {args.persona})" if args.persona else "",
                      " (preserve-intent)" if args.preserve_intent else "",
                      " (search)" if args.codex_search else ""])
    print(f"Calling {len(models)} model(s) ({mode}){extras}: {', '.join(models)}...", file=sys.stderr)

    results = call_models_parallel(models, spec, args.round, args.doc_type, args.press, args.focus, args.persona,
                                   context, args.preserve_intent, args.codex_reasoning, args.codex_search,
                                   args.timeout, bedrock_mode, bedrock_region)

    for bad in (r for r in results if r.error):
        print(f"Warning: {bad.model} returned error: {bad.error}", file=sys.stderr)
    successful = [r for r in results if not r.error]
    all_agreed = all(r.agreed for r in successful) if successful else False

    session_id = session_state.session_id if session_state else args.session
    if session_id or args.session:
        save_checkpoint(spec, args.round, session_id)

    latest_spec = next((r.spec for r in successful if r.spec), spec)
    if session_state:
        session_state.spec = latest_spec
        session_state.round = args.round + 1
        session_state.history.append({
            "round": args.round, "all_agreed": all_agreed,
            "models": [{"model": r.model, "agreed": r.agreed, "error": r.error} for r in results]})
        session_state.save()
    if args.telegram:
        print("Warning: --telegram is not part of the local build; skipping notification", file=sys.stderr)
    output_results(args, results, models, all_agreed, None, session_state)


def output_results(args, results, models: list[str], all_agreed: bool, user_feedback: Optional[str],
                   session_state) -> None:
    if args.json:
        output: dict[str, Any] = {
            "all_agreed": all_agreed, "round": args.round, "doc_type": args.doc_type, "models": models,
            "focus": args.focus, "persona": args.persona, "preserve_intent": args.preserve_intent,
            "session": session_state.session_id if session_state else args.session,
            "results": [{"model": r.model, "agreed": r.agreed, "response": r.response, "spec": r.spec,
                         "error": r.error, "input_tokens": r.input_tokens, "output_tokens": r.output_tokens,
                         "cost": r.cost} for r in results],
            "cost": {"total": cost_tracker.total_cost, "input_tokens": cost_tracker.total_input_tokens,
                     "output_tokens": cost_tracker.total_output_tokens, "by_model": cost_tracker.by_model},
        }
        if user_feedback:
            output["user_feedback"] = user_feedback
        print(json.dumps(output, indent=2))
        return
    print(f"\n=== Round {args.round} Results ({get_doc_type_name(args.doc_type)}) ===\n")
    for r in results:
        print(f"--- {r.model} ---")
        print(f"ERROR: {r.error}" if r.error else ("[AGREE]" if r.agreed else r.response))
        print()
    if all_agreed:
        print("=== ALL MODELS AGREE ===")
    else:
        ok = [r for r in results if not r.error]
        agreed = [r.model for r in ok if r.agreed]
        critiqued = [r.model for r in ok if not r.agreed]
        if agreed:
            print(f"Agreed: {', '.join(agreed)}")
        if critiqued:
            print(f"Critiqued: {', '.join(critiqued)}")
    if user_feedback:
        print("\n=== User Feedback ===")
        print(user_feedback)
    if args.show_cost:
        print(cost_tracker.summary())


def handle_export_tasks(args, models: list[str]) -> None:
    """`export-tasks`: the other `completion` call site (reference debate.py:688-736) — one model, one user
    message, max_tokens 8000, temperature 0.3, no retry; a local b200/ model runs on the same engine."""
    spec = read_spec_from_stdin()
    if not spec:
        print("Error: No spec provided via stdin", file=sys.stderr)
        sys.exit(1)
    prompt = EXPORT_TASKS_PROMPT.format(doc_type_name=get_doc_type_name(args.doc_type), spec=spec)
    try:
        kwargs = {"model": models[0], "messages": [{"role": "user", "content": prompt}], "max_tokens": 8000}
        if not is_o_series_model(models[0]):
            kwargs["temperature"] = 0.3
        tasks = extract_tasks(completion(**kwargs).choices[0].message.content)
        if args.json:
            print(json.dumps({"tasks": tasks}, indent=2))
            return
        print(f"\n=== Extracted {len(tasks)} Tasks ===\n")
        for i, task in enumerate(tasks, 1):
            print(f"{i}. [{task.get('type', 'task')}] [{task.get('priority', 'medium')}] {task.get('title', 'Untitled')}")
            if task.get("description"):
                print(f"   {task['description'][:100]}...")
            if task.get("acceptance_criteria"):
                print(f"   Acceptance criteria: {len(task['acceptance_criteria'])} items")
            print()
    except Exception as e:
        print(f"Error: {e}", file=sys.stderr)
        sys.exit(1)


def _tensor_parallel_follower() -> bool:
    """Under `torchrun` with ADVSPEC_TP=k every rank runs this CLI with the same arguments and stdin (the
    ranks of one tensor-parallel engine make identical calls); only rank 0 reports and writes files."""
    try:
        return int(os.environ.get("ADVSPEC_TP", "1")) > 1 and int(os.environ.get("RANK", "0")) != 0
    except ValueError:
        return False


def _forward_to_server(args) -> bool:
    """ADVSPEC_SERVER=<unix socket>: run this invocation in the resident server process (resident.py) so the
    engines it loaded for earlier rounds are reused; stdout, stderr and the exit code come back unchanged."""
    path = os.environ.get("ADVSPEC_SERVER")
    if not path or args.action not in ("critique", "export-tasks"):
        return False
    if __package__ in (None, ""):
        from advspec_b200 import resident
    else:
        from . import resident
    needs_stdin = not (args.action == "critique" and args.resume)
    stdin_text = sys.stdin.read() if needs_stdin else ""
    try:
        r = resident.request(path, sys.argv[1:], stdin_text)
    except OSError as e:
        print(f"Error: ADVSPEC_SERVER={path} is set but no server answers there ({e}); start one with "
              f"`python adversarial-spec_b200/resident.py serve --socket {path}` or unset ADVSPEC_SERVER",
              file=sys.stderr)
        sys.exit(2)
    sys.stdout.write(r.get("stdout", ""))
    sys.stderr.write(r.get("stderr", ""))
    sys.stdout.flush()
    sys.exit(int(r.get("code", 1)))


def main(_forward: bool = True) -> None:
    if _tensor_parallel_follower():
        sink = open(os.devnull, "w")
        sys.stdout = sys.stderr = sink
        _session.SessionState.save = lambda self: None
        globals()["save_checkpoint"] = lambda *a, **k: None
    args = create_parser().parse_args()
    if _forward and _forward_to_server(args):
        return
    if args.action == "providers":
        _providers.list_providers()
        return
    if args.action == "export-tasks":
        models = parse_models(args)
        validate_models_before_run(models, False)
        handle_export_tasks(args, models)
        return
    if args.action != "critique":
        print(f"Error: '{args.action}' is outside the local engine's scope (it is config or a single remote "
              f"call with no fan-out); use the reference CLI for it.", file=sys.stderr)
        sys.exit(2)
    if args.profile:
        print("Warning: --profile is not part of the local build; ignoring", file=sys.stderr)
    models = parse_models(args)
    context = load_context_files(args.context) if args.context else None
    validate_models_before_run(models, False)
    spec, session_state, models = load_or_resume_session(args, models)
    run_critique(args, spec, models, session_state, context, False, None)


if __name__ == "__main__":
    main()
