#!/usr/bin/env python3
"""Generates tests/golden/* IN THE BUILD CONTAINER (needs /root/reference and transformers).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Two kinds of fixture:
 1. reference_cli_cases.json — the UNMODIFIED reference CLI
    (/root/reference/skills/adversarial-spec/scripts/debate.py) run as a
    subprocess with a canned ``litellm`` stub on PYTHONPATH; records stdout,
    stderr markers and exit code per case.  Pins the fan-out / parsing / output
    contract for tests/test_reference_pin.py.
 2. hf_logits_<model>.npz — transformers-5.5.0 fp32 logits of the seeded
    random-init tiny models on fixed token ids.  Pins the numeric oracle itself
    (the reference has no numeric fixture: SURVEY.md §8(c)).
Re-run:  python oracle/make_golden.py
"""

from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
REF_SCRIPTS = Path("/root/reference/skills/adversarial-spec/scripts")
GOLDEN = ROOT / "tests" / "golden"

CANNED_STUB = '''
import json, os, time
suppress_debug_info = False
_CASE = json.loads(os.environ["ADVSPEC_CANNED"])
class _O:
    def __init__(self, **k): self.__dict__.update(k)
def completion(**kw):
    spec = _CASE["responses"][kw["model"]]
    time.sleep(spec.get("delay", 0.0))
    if spec.get("raise"):
        raise RuntimeError(spec["raise"])
    usage = None if spec.get("no_usage") else _O(prompt_tokens=spec["in"], completion_tokens=spec["out"])
    return _O(choices=[_O(message=_O(content=spec["content"]))], usage=usage)
'''

SPEC_TEXT = "# Payments API\n\nThe service must store data.\n\n## Limits\nrate limit 100 qps"

CASES = [
    {"name": "two_models_one_agrees_json",
     "argv": ["critique", "--models", "fake/a,fake/b", "--doc-type", "tech", "--json"],
     "responses": {"fake/a": {"content": "Looks complete.\n[AGREE]\n[SPEC]final text[/SPEC]", "in": 120, "out": 30, "delay": 0.0},
                   "fake/b": {"content": "Missing error model.\n[SPEC]\nrevised spec body\n[/SPEC]", "in": 121, "out": 44, "delay": 0.3}}},
    {"name": "all_agree_text",
     "argv": ["critique", "--models", "fake/a,fake/b", "--doc-type", "prd", "--round", "3"],
     "responses": {"fake/a": {"content": "[AGREE]\n[SPEC]s[/SPEC]", "in": 10, "out": 5, "delay": 0.0},
                   "fake/b": {"content": "ok [AGREE] [SPEC]s2[/SPEC]", "in": 11, "out": 6, "delay": 0.3}}},
    {"name": "malformed_reply_and_no_usage_json",
     "argv": ["critique", "--models", "fake/a", "--json", "--focus", "security", "--preserve-intent"],
     "responses": {"fake/a": {"content": "just a critique without tags", "no_usage": True}}},
    {"name": "one_model_fails_text_show_cost",
     "argv": ["critique", "--models", "fake/a,fake/b", "--show-cost", "--persona", "security engineer"],
     "responses": {"fake/a": {"raise": "boom: provider exploded"},
                   "fake/b": {"content": "crit [SPEC]x[/SPEC]", "in": 1000, "out": 2000, "delay": 0.1}}},
    {"name": "empty_stdin", "argv": ["critique", "--models", "fake/a"], "stdin": "   \n",
     "responses": {"fake/a": {"content": "x", "in": 1, "out": 1}}},
    {"name": "missing_key_exit2", "argv": ["critique", "--models", "gpt-4o"],
     "responses": {"gpt-4o": {"content": "x", "in": 1, "out": 1}}},
    {"name": "export_tasks_json", "argv": ["export-tasks", "--models", "fake/a", "--doc-type", "prd", "--json"],
     "responses": {"fake/a": {"content": "intro\n[TASK]\ntitle: Add login\ntype: user-story\npriority: high\ndescription: Users sign in\nwith email.\nacceptance_criteria:\n- form exists\n- errors shown\n[/TASK]\n[TASK]\ntype: bug\n[/TASK]\n[TASK]\ntitle: Spike cache\nacceptance_criteria:\n- measured\npriority: low\n[/TASK] trailing", "in": 5, "out": 5}}},
    {"name": "export_tasks_text", "argv": ["export-tasks", "--models", "fake/a"],
     "responses": {"fake/a": {"content": "[TASK]\ntitle: Rate limit the API\ntype: task\npriority: medium\ndescription: " + "x" * 130 + "\nacceptance_criteria:\n- 429 returned\n[/TASK]", "in": 5, "out": 5}}},
    {"name": "export_tasks_model_error", "argv": ["export-tasks", "--models", "fake/a"],
     "responses": {"fake/a": {"raise": "provider down"}}},
    {"name": "press_round_json",
     "argv": ["critique", "--models", "fake/a", "--press", "--json", "--round", "2"],
     "responses": {"fake/a": {"content": "verified [AGREE]\n[SPEC] kept [/SPEC]", "in": 50, "out": 9}}},
]


# Multi-step flows through ONE home + working directory: session create -> resume, context files.
# The stub also records what the reference puts on the wire (wire.jsonl) so the envelope of resumed,
# pressed and context-carrying rounds is pinned too.
FLOW_STUB = CANNED_STUB + '''
_orig_completion = completion
def completion(**kw):
    with open(os.path.join(os.environ["HOME"], "wire.jsonl"), "a") as f:
        f.write(json.dumps({"model": kw["model"], "messages": kw["messages"],
                            "max_tokens": kw.get("max_tokens"), "temperature": kw.get("temperature")}) + "\\n")
    return _orig_completion(**kw)
'''

FLOWS = [
    {"name": "session_create_then_resume",
     "steps": [
         {"argv": ["critique", "--models", "fake/a,fake/b", "--session", "s1", "--json", "--focus", "security"],
          "responses": {"fake/a": {"content": "needs work\n[SPEC]\nrevised once\n[/SPEC]", "in": 40, "out": 8, "delay": 0.0},
                        "fake/b": {"content": "fine [AGREE]", "in": 41, "out": 3, "delay": 0.3}}},
         {"argv": ["critique", "--resume", "s1", "--models", "fake/a,fake/b", "--show-cost"], "stdin": "",
          "responses": {"fake/a": {"content": "[AGREE]\n[SPEC]final[/SPEC]", "in": 12, "out": 4, "delay": 0.0},
                        "fake/b": {"content": "[AGREE]", "in": 13, "out": 2, "delay": 0.3}}},
     ]},
    {"name": "resume_without_models_is_refused",  # model selection runs before the session is read
     "steps": [{"argv": ["critique", "--resume", "s1"], "stdin": "",
                "responses": {"fake/a": {"content": "x", "in": 1, "out": 1}}}]},
    {"name": "resume_unknown_session",
     "steps": [{"argv": ["critique", "--resume", "nope", "--models", "fake/a"], "stdin": "",
                "responses": {"fake/a": {"content": "x", "in": 1, "out": 1}}}]},
    {"name": "context_files_and_checkpoint",
     "files": {"api.md": "# API\nGET /v1/things\n", "schema.sql": "create table t (id int);\n"},
     "steps": [{"argv": ["critique", "--models", "fake/a", "--context", "api.md", "--context", "schema.sql",
                         "--context", "missing.txt", "--session", "ctx", "--round", "2", "--doc-type", "prd"],
                "responses": {"fake/a": {"content": "crit\n[SPEC]with context[/SPEC]", "in": 70, "out": 9}}}]},
]


def _snapshot(td: str) -> dict:
    """Files the flow left behind, with timestamps and the temporary directory normalised."""
    import re
    out = {}
    for pth in sorted(Path(td).rglob("*")):
        rel = pth.relative_to(td).as_posix()
        if not pth.is_file() or rel == "litellm.py" or "__pycache__" in rel:
            continue
        txt = pth.read_text()
        txt = re.sub(r'"(created_at|updated_at|timestamp)": "[^"]*"', lambda m: '"%s": "<ts>"' % m.group(1), txt)
        txt = txt.replace(td, "<DIR>")
        out[rel] = txt
    return out


def run_reference_flow(flow: dict) -> dict:
    steps = []
    with tempfile.TemporaryDirectory() as td:
        (Path(td) / "litellm.py").write_text(FLOW_STUB)
        for name, content in flow.get("files", {}).items():
            (Path(td) / name).write_text(content)
        for step in flow["steps"]:
            env = {k: v for k, v in os.environ.items() if not k.endswith("_API_KEY")}
            env.update(PYTHONPATH=td, ADVSPEC_CANNED=json.dumps(step), HOME=td)
            p = subprocess.run([sys.executable, str(REF_SCRIPTS / "debate.py"), *step["argv"]],
                               input=step.get("stdin", SPEC_TEXT), capture_output=True, text=True, env=env, cwd=td)
            steps.append({"stdout": p.stdout.replace(td, "<DIR>"), "stderr": p.stderr.replace(td, "<DIR>"),
                          "returncode": p.returncode})
        snap = _snapshot(td)
    return {"steps": steps, "files": snap}


def make_reference_flows() -> None:
    out = []
    for flow in FLOWS:
        got = run_reference_flow(flow)
        for st in flow["steps"]:
            st.setdefault("stdin", SPEC_TEXT)
        out.append({**flow, "expected": got})
        print(f"flow {flow['name']}: rcs={[s['returncode'] for s in got['steps']]} files={sorted(got['files'])}")
    (GOLDEN / "reference_cli_flows.json").write_text(json.dumps(out, indent=1))


def run_reference_cli(case: dict) -> dict:
    with tempfile.TemporaryDirectory() as td:
        (Path(td) / "litellm.py").write_text(CANNED_STUB)
        env = {k: v for k, v in os.environ.items() if not k.endswith("_API_KEY")}
        env["PYTHONPATH"] = td
        env["ADVSPEC_CANNED"] = json.dumps(case)
        env["HOME"] = td  # sessions/config stay out of the real home
        p = subprocess.run([sys.executable, str(REF_SCRIPTS / "debate.py"), *case["argv"]],
                           input=case.get("stdin", SPEC_TEXT), capture_output=True, text=True, env=env, cwd=td)
    return {"stdout": p.stdout, "stderr": p.stderr, "returncode": p.returncode}


def make_reference_cases() -> None:
    out = []
    for case in CASES:
        got = run_reference_cli(case)
        out.append({**case, "stdin": case.get("stdin", SPEC_TEXT), "expected": got})
        print(f"{case['name']}: rc={got['returncode']} stdout={len(got['stdout'])}B")
    (GOLDEN / "reference_cli_cases.json").write_text(json.dumps(out, indent=1))


def make_message_fixtures() -> None:
    """What the reference puts on the wire (the token stream the engine prefills): sha256 and
    length of system/user messages for several flag combinations, captured at seam B1."""
    import hashlib

    sys.path.insert(0, str(REF_SCRIPTS))
    with tempfile.TemporaryDirectory() as td:
        (Path(td) / "litellm.py").write_text("suppress_debug_info=False\ndef completion(**kw):\n    raise RuntimeError('unused')\n")
        sys.path.insert(0, td)
        import models as ref_models  # the reference's module
        captured = {}

        def fake(**kw):
            captured.update(kw)
            raise RuntimeError("captured")

        ref_models.completion = fake
        ref_models.time.sleep = lambda s: None
        rows = []
        combos = [dict(doc_type="prd"), dict(doc_type="tech"), dict(doc_type="tech", press=True),
                  dict(doc_type="tech", focus="security"), dict(doc_type="prd", focus="latency budget"),
                  dict(doc_type="tech", persona="oncall-engineer"), dict(doc_type="tech", persona="Space Pirate"),
                  dict(doc_type="prd", preserve_intent=True, context="## Additional Context\nctx")]
        for kw in combos:
            captured.clear()
            ref_models.call_single_model("fake/a", SPEC_TEXT, 4, **kw)
            sysm, usr = captured["messages"][0]["content"], captured["messages"][1]["content"]
            rows.append({"kwargs": kw, "system_sha256": hashlib.sha256(sysm.encode()).hexdigest(),
                         "system_len": len(sysm), "user_sha256": hashlib.sha256(usr.encode()).hexdigest(),
                         "user_len": len(usr), "max_tokens": captured["max_tokens"],
                         "temperature": captured.get("temperature"), "roles": [m["role"] for m in captured["messages"]],
                         "user_prefix": usr[:60], "user_suffix": usr[-60:]})
    (GOLDEN / "reference_messages.json").write_text(json.dumps({"spec": SPEC_TEXT, "round": 4, "rows": rows}, indent=1))
    print(f"reference_messages.json: {len(rows)} rows")


HF_GOLDEN_NAMES = ["tiny-llama", "tiny-llama-128", "tiny-qwen2", "tiny-gemma", "tiny-phi3", "tiny-gemma256",
                   "tiny-mistral"]


def make_hf_logits(only=None) -> None:
    import numpy as np

    import advspec_loader

    advspec_loader.load()
    from advspec_b200 import model_spec
    from oracle import hf_oracle

    import transformers

    for name in (only or HF_GOLDEN_NAMES):
        spec = model_spec.resolve(name)
        model = hf_oracle.build_hf_model(spec, 1234)
        toks = np.random.default_rng(7).integers(0, spec.vocab_size, 24)
        logits = hf_oracle.hf_logits(model, toks.tolist())
        blob = hf_oracle.export_blob(spec, model)
        import hashlib

        np.savez_compressed(GOLDEN / f"hf_logits_{name}.npz", tokens=toks.astype(np.int32),
                            logits=logits.astype(np.float32), seed=np.int64(1234),
                            inv_freq=hf_oracle.rope_inv_freq(model),
                            blob_sha256=np.frombuffer(hashlib.sha256(blob.tobytes()).digest(), dtype=np.uint8),
                            transformers=np.bytes_(transformers.__version__))
        print(f"hf_logits_{name}.npz: logits {logits.shape}")


if __name__ == "__main__":
    GOLDEN.mkdir(parents=True, exist_ok=True)
    if len(sys.argv) > 2 and sys.argv[1] == "--hf-only":  # add a model's logits without touching the rest
        make_hf_logits(sys.argv[2:])
        sys.exit(0)
    if REF_SCRIPTS.exists():
        make_reference_cases()
        make_reference_flows()
        make_message_fixtures()
    else:
        print("no /root/reference here: reference fixtures left untouched")
    make_hf_logits()
