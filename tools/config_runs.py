"""BASELINE configs 3-5 as far as one GPU carries them, through the public host API
(`models.call_models_parallel`): heterogeneous 5-model panel at 8K, a 4-replica panel at 16K over
two rounds, and a 32K-token prompt.  Prints one JSON object per run."""
import json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import advspec_loader; advspec_loader.load()
from advspec_b200 import models, runtime, model_spec
from advspec_b200.tokenizer import SyntheticTokenizer, generate_spec

if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("ADVSPEC_TP"):
    os.environ["ADVSPEC_DEVICES"] = os.environ.get("LOCAL_RANK", "0")  # one rank of a tensor-parallel group
os.environ.setdefault("ADVSPEC_DEVICES", "0")
gen = int(os.environ.get("CFG_GEN", "32"))
os.environ["ADVSPEC_MAX_NEW_TOKENS"] = str(gen)
which = sys.argv[1:] or ["hetero8k", "replica16k", "long32k"]
devnull = open(os.devnull, "w")


def run(name, panel, spec_tokens, doc_type, rounds=1):
    tok = SyntheticTokenizer(32000)
    doc = generate_spec(tok, spec_tokens, seed=11).strip()
    out = {"run": name, "panel": panel, "spec_tokens": spec_tokens, "new_tokens": gen, "rounds": []}
    for r in range(1, rounds + 1):
        err, sys.stderr = sys.stderr, devnull
        t0 = time.perf_counter()
        try:
            res = models.call_models_parallel(panel, doc, r, doc_type)
        finally:
            sys.stderr = err
        wall = time.perf_counter() - t0
        rows = [{"model": x.model, "in": x.input_tokens, "out": x.output_tokens, "error": x.error} for x in res]
        eng = {f"{k[0]}@{k[1]}": {"prefill_ms": round(v.engine.timing().prefill_ms, 2),
                                  "decode_ms_per_step": round(v.engine.timing().decode_ms / max(v.engine.timing().decode_steps, 1), 3),
                                  "batch": v.engine.timing().decode_batch}
               for k, v in runtime.POOL._engines.items()}
        out["rounds"].append({"round": r, "wall_s": round(wall, 3), "tokens_per_s": round(sum(x.output_tokens for x in res) / wall, 1),
                              "results": rows, "engines": eng})
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(out), flush=True)
    runtime.POOL.close()


if "hetero8k" in which:
    run("config3: heterogeneous panel, 8K spec, one GPU, engines side by side",
        ["b200/llama-3-8b", "b200/mistral-7b", "b200/qwen2-7b", "b200/phi-3-mini", "b200/gemma-7b"], 8192, "tech")
if "replica16k" in which:
    run("config4: 4 replicas, 16K tech spec, 2 of the 6 rounds (weights stay resident between rounds)",
        ["b200/llama-3-8b"] * 4, 16384, "tech", rounds=2)
if "long32k" in which:
    run("32K-token spec, one Llama-3-8B opponent (4 prefill chunks)", ["b200/llama-3-8b"], 32768, "tech")
if "tp" in which:
    # config 5 through the public API: torchrun, one process per GPU, ADVSPEC_TP = WORLD_SIZE; every rank
    # makes the same call and gets the same critique
    m = os.environ.get("CFG_TP_MODEL", "llama-3-8b")
    run(f"config5-style: one {m} opponent tensor-parallel over {os.environ.get('WORLD_SIZE', '1')} GPUs",
        [f"b200/{m}"], int(os.environ.get("CFG_TP_SPEC", "4096")), "tech")
if "converge16k" in which:
    # config 4 the way the reference drives it — one `debate.py critique` per round, `--session` then
    # `--resume` — but in ONE process, so the engine (weights, workspaces) stays resident across rounds
    import contextlib, io, tempfile
    from advspec_b200 import debate, session
    rounds = int(os.environ.get("CFG_ROUNDS", "6"))
    n_tok = int(os.environ.get("CFG_SPEC", "16384"))
    doc = generate_spec(SyntheticTokenizer(32000), n_tok, seed=11).strip()
    td = tempfile.mkdtemp()
    session.SESSIONS_DIR = Path(td) / ".config" / "adversarial-spec" / "sessions"
    session.CHECKPOINTS_DIR = Path(td) / ".adversarial-spec-checkpoints"
    panel = ",".join(["b200/llama-3-8b"] * 4)
    out = {"run": f"config4: 4 replicas x {rounds} rounds, {n_tok}-token tech spec, debate.py critique --session/--resume in one process",
           "new_tokens": gen, "rounds": []}
    for r in range(1, rounds + 1):
        argv = ["debate.py", "critique", "--models", panel, "--doc-type", "tech", "--json"] + \
               (["--session", "cfg4"] if r == 1 else ["--resume", "cfg4"])
        sys.argv, sys.stdin = argv, io.StringIO(doc if r == 1 else "")
        so, se = io.StringIO(), io.StringIO()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(so), contextlib.redirect_stderr(se):
            try:
                debate.main()
                rc = 0
            except SystemExit as e:
                rc = e.code or 0
        wall = time.perf_counter() - t0
        res = json.loads(so.getvalue()) if rc == 0 else {}
        eng = {f"{k[0]}@{k[1]}": {"prefill_ms": round(v.engine.timing().prefill_ms, 2),
                                  "decode_ms_per_step": round(v.engine.timing().decode_ms / max(v.engine.timing().decode_steps, 1), 3),
                                  "batch": v.engine.timing().decode_batch} for k, v in runtime.POOL._engines.items()}
        toks = sum(x["output_tokens"] for x in res.get("results", []))
        out["rounds"].append({"round": res.get("round"), "rc": rc, "wall_s": round(wall, 3), "output_tokens": toks,
                              "tokens_per_s": round(toks / wall, 1), "all_agreed": res.get("all_agreed"),
                              "input_tokens": [x["input_tokens"] for x in res.get("results", [])], "engines": eng,
                              "stderr_first_line": se.getvalue().splitlines()[0] if se.getvalue() else ""})
    out["session_file_rounds"] = json.loads((session.SESSIONS_DIR / "cfg4.json").read_text())["round"]
    out["checkpoints"] = sorted(p.name for p in session.CHECKPOINTS_DIR.glob("*.md"))
    print(json.dumps(out), flush=True)
    runtime.POOL.close()
