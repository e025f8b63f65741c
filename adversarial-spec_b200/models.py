"""Fan-out of one critique round — the host-side mirror of the reference's hot path.

Same names, argument order, return types and error behaviour as
skills/adversarial-spec/scripts/models.py:
  ModelResponse :67-78, CostTracker :81-123, cost_tracker :127,
  load_context_files :130-146, detect_agreement :149-151, extract_spec :154-160,
  call_single_model :457-678 (litellm branch), call_models_parallel :681-722.
What differs is below the seam: opponents named ``b200/<model>`` are not sent
one by one to a provider; ``call_models_parallel`` sees the whole panel, groups
same-weight opponents, and runs ONE shared-prefix prefill + batched decode per
group on the local GPU (runtime.run_round).  Everything the caller can observe
is kept: results arrive in completion order, a failing opponent is retried 3x
with 1 s / 2 s back-off and then reported in ``.error`` without sinking the
others, ``[AGREE]`` / ``[SPEC]`` parsing and cost accounting are unchanged.
"""

from __future__ import annotations

import concurrent.futures
import os
import sys
import time
from dataclasses import dataclass, field
from pathlib import Path
from typing import Optional

from . import runtime
from .completion import completion  # seam B1; tests patch ``models.completion`` like the reference's
from .envelope import build_messages, build_panel_messages
from .model_spec import is_local_model, split_persona
from .providers import DEFAULT_CODEX_REASONING, DEFAULT_COST, MODEL_COSTS

MAX_RETRIES = 3
RETRY_BASE_DELAY = 1.0  # seconds


def is_o_series_model(model: str) -> bool:
    m = model.lower()
    return m.startswith("o1") or "/o1" in m or "-o1" in m


@dataclass
class ModelResponse:
    model: str
    response: str
    agreed: bool
    spec: Optional[str]
    error: Optional[str] = None
    input_tokens: int = 0
    output_tokens: int = 0
    cost: float = 0.0


@dataclass
class CostTracker:
    total_input_tokens: int = 0
    total_output_tokens: int = 0
    total_cost: float = 0.0
    by_model: dict = field(default_factory=dict)

    def add(self, model: str, input_tokens: int, output_tokens: int) -> float:
        rate = MODEL_COSTS.get(model) or MODEL_COSTS.get(split_persona(model)[0], DEFAULT_COST)
        cost = input_tokens / 1_000_000 * rate["input"] + output_tokens / 1_000_000 * rate["output"]
        self.total_input_tokens += input_tokens
        self.total_output_tokens += output_tokens
        self.total_cost += cost
        row = self.by_model.setdefault(model, {"input_tokens": 0, "output_tokens": 0, "cost": 0.0})
        row["input_tokens"] += input_tokens
        row["output_tokens"] += output_tokens
        row["cost"] += cost
        return cost

    def summary(self) -> str:
        lines = ["", "=== Cost Summary ===",
                 f"Total tokens: {self.total_input_tokens:,} in / {self.total_output_tokens:,} out",
                 f"Total cost: ${self.total_cost:.4f}"]
        if len(self.by_model) > 1:
            lines += ["", "By model:"]
            for model, d in self.by_model.items():
                lines.append(f"  {model}: ${d['cost']:.4f} ({d['input_tokens']:,} in / {d['output_tokens']:,} out)")
        return "\n".join(lines)


cost_tracker = CostTracker()


def load_context_files(context_paths: list[str]) -> str:
    if not context_paths:
        return ""
    sections = []
    for path in context_paths:
        try:
            sections.append(f"### Context: {path}\n```\n{Path(path).read_text()}\n```")
        except Exception as e:
            sections.append(f"### Context: {path}\n[Error loading file: {e}]")
    return ("## Additional Context\nThe following documents are provided as context:\n\n"
            + "\n\n".join(sections))


def detect_agreement(response: str) -> bool:
    return "[AGREE]" in response


def extract_spec(response: str) -> Optional[str]:
    if "[SPEC]" not in response or "[/SPEC]" not in response:
        return None
    start = response.find("[SPEC]") + len("[SPEC]")
    return response[start: response.find("[/SPEC]")].strip()


_TASK_KEYS = ("title:", "type:", "priority:", "description:", "acceptance_criteria:")


def extract_tasks(response: str) -> list[dict]:
    """[TASK]...[/TASK] blocks -> dicts (reference parser: models.py:163-247, quirks included: a key line
    closes the previous field; a field with several lines is joined with newlines; acceptance_criteria is
    a list only when it is the last field of the block; blocks without a title are dropped)."""

    def close(values: list[str]):
        return "\n".join(values).strip() if len(values) > 1 else (values[0] if values else "")

    tasks = []
    for block in response.split("[TASK]")[1:]:
        if "[/TASK]" not in block:
            continue
        task: dict = {}
        key: Optional[str] = None
        values: list[str] = []
        for raw in block.split("[/TASK]")[0].strip().split("\n"):
            line = raw.strip()
            head = next((k for k in _TASK_KEYS if line.startswith(k)), None)
            if head is not None:
                if key:
                    task[key] = close(values)
                key = head[:-1]
                values = [] if key == "acceptance_criteria" else [line[len(head):].strip()]
            elif line.startswith("- ") and key == "acceptance_criteria":
                values.append(line[2:])
            elif key:
                values.append(line)
        if key:
            task[key] = values if key == "acceptance_criteria" else "\n".join(values).strip()
        if task.get("title"):
            tasks.append(task)
    return tasks


def _finish(model: str, content: str, input_tokens: int, output_tokens: int) -> ModelResponse:
    agreed = "[AGREE]" in content
    extracted = extract_spec(content)
    if not agreed and not extracted:
        print(f"Warning: {model} provided critique but no [SPEC] tags found. Response may be malformed.",
              file=sys.stderr)
    cost = cost_tracker.add(model, input_tokens, output_tokens)
    return ModelResponse(model=model, response=content, agreed=agreed, spec=extracted,
                         input_tokens=input_tokens, output_tokens=output_tokens, cost=cost)


def _report_failure(model: str, attempt: int, last_error: str, sleep: bool = True) -> float:
    """The reference's retry messages (models.py:663-676).  Returns the back-off of this attempt (0 after
    the last); sleeps it unless the caller does so once for a whole group of opponents."""
    if attempt < MAX_RETRIES - 1:
        delay = RETRY_BASE_DELAY * (2 ** attempt)
        print(f"Warning: {model} failed (attempt {attempt + 1}/{MAX_RETRIES}): {last_error}. "
              f"Retrying in {delay:.1f}s...", file=sys.stderr)
        if sleep:
            time.sleep(delay)
        return delay
    print(f"Error: {model} failed after {MAX_RETRIES} attempts: {last_error}", file=sys.stderr)
    return 0.0


def call_single_model(model: str, spec: str, round_num: int, doc_type: str, press: bool = False,
                      focus: Optional[str] = None, persona: Optional[str] = None,
                      context: Optional[str] = None, preserve_intent: bool = False,
                      codex_reasoning: str = DEFAULT_CODEX_REASONING, codex_search: bool = False,
                      timeout: int = 600, bedrock_mode: bool = False,
                      bedrock_region: Optional[str] = None) -> ModelResponse:
    """One opponent through seam B1 (``completion``), with the reference's retry policy."""
    actual_model = model
    if bedrock_mode:
        if bedrock_region:
            os.environ["AWS_REGION"] = bedrock_region
        if not model.startswith("bedrock/"):
            actual_model = f"bedrock/{model}"
    if model.startswith("codex/") or model.startswith("gemini-cli/"):
        # subscription CLIs are remote models: nothing to run locally (SURVEY.md §2.1 out of scope)
        return ModelResponse(model=model, response="", agreed=False, spec=None,
                             error="CLI-tool backends (codex/, gemini-cli/) are not part of the local engine")
    system_prompt, user_message = build_messages(spec, round_num, doc_type, press, focus, persona, context,
                                                 preserve_intent)
    last_error = None
    for attempt in range(MAX_RETRIES):
        try:
            kwargs = {
                "model": actual_model,
                "messages": [{"role": "system", "content": system_prompt},
                             {"role": "user", "content": user_message}],
                "max_tokens": 8000,
                "timeout": timeout,
            }
            if not is_o_series_model(actual_model):
                kwargs["temperature"] = 0.7
            response = completion(**kwargs)
            content = response.choices[0].message.content
            in_tok = response.usage.prompt_tokens if response.usage else 0
            out_tok = response.usage.completion_tokens if response.usage else 0
            return _finish(model, content, in_tok, out_tok)
        except Exception as e:
            last_error = str(e)
            if bedrock_mode:
                if "AccessDeniedException" in last_error:
                    last_error = f"Model not enabled in your Bedrock account: {model}"
                elif "ValidationException" in last_error:
                    last_error = f"Invalid Bedrock model ID: {model}"
            _report_failure(model, attempt, last_error)
    return ModelResponse(model=model, response="", agreed=False, spec=None, error=last_error)


def _call_local_panel(local: list[tuple[int, str]], spec: str, round_num: int, doc_type: str, press: bool,
                      focus, persona, context, preserve_intent: bool, timeout: int = 600) -> list[ModelResponse]:
    """All local opponents of the round in one engine pass per same-weight group; opponents whose group
    raises (or whose name is unknown) are retried together on the reference's schedule — one back-off per
    attempt, as the reference's threads sleep in parallel — then reported per opponent.  `timeout` bounds
    each attempt like the per-call timeout of the reference (models.py:621)."""
    names = [m for _, m in local]
    # `b200/<model>@<persona>` gives that opponent a persona of its own; the others keep
    # the panel's --persona.  With one persona for everyone this is exactly build_messages.
    personas = [split_persona(m)[1] or persona for m in names]
    system_prompt, user_messages = build_panel_messages(spec, round_num, doc_type, press, focus, personas,
                                                        context, preserve_intent)
    same = len(set(user_messages)) <= 1
    seeds = [runtime.opponent_seed(round_num, i) for i, _ in local]
    done: dict[int, ModelResponse] = {}
    pending = list(range(len(local)))
    for attempt in range(MAX_RETRIES):
        one = concurrent.futures.ThreadPoolExecutor(max_workers=1)
        fut = one.submit(runtime.run_round, [names[j] for j in pending], system_prompt,
                         user_messages[0] if same else [user_messages[j] for j in pending],
                         [seeds[j] for j in pending], 8000, 0.7)
        try:
            outs = fut.result(timeout=timeout if timeout and timeout > 0 else None)
        except concurrent.futures.TimeoutError:
            # a GPU round cannot be cancelled: it finishes in the background (holding its engine lease)
            outs = [TimeoutError(f"local engine round timed out after {timeout}s")] * len(pending)
        except Exception as ex:  # run_round reports per opponent; anything else fails the attempt as a whole
            outs = [ex] * len(pending)
        one.shutdown(wait=False)
        failed: list[int] = []
        for j, out in zip(pending, outs):
            if isinstance(out, Exception):
                failed.append(j)
                done[j] = ModelResponse(model=names[j], response="", agreed=False, spec=None, error=str(out))
            else:
                done[j] = _finish(names[j], out.text, out.prompt_tokens, out.completion_tokens)
        if not failed:
            break
        delay = max(_report_failure(names[j], attempt, done[j].error or "", sleep=False) for j in failed)
        if delay > 0:
            time.sleep(delay)
        pending = failed
    return [done[j] for j in range(len(local))]


def call_models_parallel(models: list[str], spec: str, round_num: int, doc_type: str, press: bool = False,
                         focus: Optional[str] = None, persona: Optional[str] = None,
                         context: Optional[str] = None, preserve_intent: bool = False,
                         codex_reasoning: str = DEFAULT_CODEX_REASONING, codex_search: bool = False,
                         timeout: int = 600, bedrock_mode: bool = False,
                         bedrock_region: Optional[str] = None) -> list[ModelResponse]:
    """Seam B2.  Local opponents run as one batched engine round; any other model string goes
    through ``call_single_model`` on its own thread exactly as in the reference.  Results are
    returned in completion order (the reference uses ``as_completed``, models.py:720-721)."""
    local = [(i, m) for i, m in enumerate(models) if is_local_model(m) and not bedrock_mode]
    remote = [(i, m) for i, m in enumerate(models) if not (is_local_model(m) and not bedrock_mode)]
    results: list[ModelResponse] = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, len(remote) + 1)) as pool:
        futures = []
        if local:
            futures.append(pool.submit(_call_local_panel, local, spec, round_num, doc_type, press, focus,
                                       persona, context, preserve_intent, timeout))
        for _, m in remote:
            futures.append(pool.submit(call_single_model, m, spec, round_num, doc_type, press, focus, persona,
                                       context, preserve_intent, codex_reasoning, codex_search, timeout,
                                       bedrock_mode, bedrock_region))
        for fut in concurrent.futures.as_completed(futures):
            r = fut.result()
            results.extend(r if isinstance(r, list) else [r])
    return results
