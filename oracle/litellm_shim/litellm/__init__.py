"""A module NAMED ``litellm`` whose ``completion`` runs a HuggingFace model on the host CPU.

TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/__init__.py).  Put this
directory on PYTHONPATH and the reference's own, unmodified ``debate.py`` /
``models.py`` import it in place of the real litellm (pinned ==1.80.13, not
installed offline, requirements.txt:1) — that is how BASELINE.md's CPU baseline
"debate.py -> litellm -> local HF model on CPU" is realised, and how
oracle/make_golden.py produced tests/golden/reference_cli_*.json.

Only what the reference touches exists: ``completion(**kwargs)`` with
``model, messages, max_tokens, timeout, temperature`` (models.py:614-628) and
``suppress_debug_info`` (models.py:22).  Same synthetic tokenizer, chat template,
weights seed and ADVSPEC_MAX_NEW_TOKENS cap as the GPU engine.
"""

from __future__ import annotations

import os
import sys
import threading
import time
from pathlib import Path
from types import SimpleNamespace

_ROOT = Path(__file__).resolve().parents[3]
if str(_ROOT) not in sys.path:
    sys.path.insert(0, str(_ROOT))

suppress_debug_info = False

_models: dict = {}
_models_mu = threading.Lock()
_tls = threading.local()
LAST_TIMINGS: list = []  # one dict per completion call (bench.py reads and clears it)
_timings_mu = threading.Lock()


def _get_model(name: str):
    import torch

    import advspec_loader

    advspec_loader.load()
    from advspec_b200.model_spec import resolve
    from advspec_b200.tokenizer import SyntheticTokenizer
    from oracle import hf_oracle

    with _models_mu:
        if name not in _models and "gpt2" in name.lower():
            # BASELINE configs[0]: the CPU-only plumbing case, GPT-2 124M random-init (no engine involved);
            # n_positions is enlarged so the prompt envelope + stub spec + generation fit
            import transformers as tf
            from types import SimpleNamespace

            torch.manual_seed(int(os.environ.get("ADVSPEC_WEIGHT_SEED", "0")))
            cfg = tf.GPT2Config(n_positions=4096)
            model = tf.AutoModelForCausalLM.from_config(cfg).float().eval()
            for layer in model.transformer.h:
                layer.register_forward_pre_hook(_pre_hook)
                layer.register_forward_hook(_post_hook)
            _models[name] = (SimpleNamespace(n_layers=cfg.n_layer, vocab_size=cfg.vocab_size), model,
                             SyntheticTokenizer(cfg.vocab_size))
        if name not in _models:
            spec = resolve(name)
            seed = int(os.environ.get("ADVSPEC_WEIGHT_SEED", "0"))
            if os.environ.get("ADVSPEC_CPU_FAST_INIT"):
                # timing-only weights: values do not matter, skip the slow normal_ init
                import transformers as tf
                from transformers.initialization import no_init_weights

                with no_init_weights():
                    model = tf.AutoModelForCausalLM.from_config(hf_oracle.hf_config(spec),
                                                                attn_implementation="eager").float().eval()
                g = torch.Generator().manual_seed(seed)
                with torch.no_grad():
                    for p in model.parameters():
                        if p.dim() == 1:
                            p.fill_(1.0)
                        else:
                            p.uniform_(-0.03, 0.03, generator=g)
            else:
                model = hf_oracle.build_hf_model(spec, seed)
            for layer in model.model.layers:
                layer.register_forward_pre_hook(_pre_hook)
                layer.register_forward_hook(_post_hook)
            _models[name] = (spec, model, SyntheticTokenizer(spec.vocab_size))
        return _models[name]


def _pre_hook(module, args):
    _tls.t_layer0 = time.perf_counter()


def _post_hook(module, args, output):
    _tls.layer_s = getattr(_tls, "layer_s", 0.0) + (time.perf_counter() - _tls.t_layer0)


def completion(*, model: str, messages: list, max_tokens: int = 8000, timeout=None, temperature: float = 1.0,
               **_ignored):
    import torch

    spec, hf, tok = _get_model(model)
    from advspec_b200.tokenizer import render_chat

    system = "\n".join(m.get("content", "") for m in messages if m.get("role") == "system")
    user = "\n".join(m.get("content", "") for m in messages if m.get("role") != "system")
    ids = tok.encode(render_chat(system, user), bos=True)
    cap = int(os.environ.get("ADVSPEC_MAX_NEW_TOKENS", "0"))
    n_new = max(1, min(max_tokens, cap) if cap > 0 else max_tokens)
    gen = torch.Generator().manual_seed(int(os.environ.get("ADVSPEC_SEED", "0")) + threading.get_ident() % 9973)
    out: list[int] = []
    with torch.no_grad():
        _tls.layer_s = 0.0
        t0 = time.perf_counter()
        r = hf(input_ids=torch.tensor([ids], dtype=torch.long), use_cache=True, logits_to_keep=1)  # as HF generate does
        t_prefill = time.perf_counter() - t0
        layer_prefill = _tls.layer_s
        _tls.layer_s = 0.0
        t1 = time.perf_counter()
        past = r.past_key_values
        logits = r.logits[0, -1]
        for _ in range(n_new):
            if temperature and temperature > 0:
                nxt = int(torch.multinomial(torch.softmax(logits.float() / temperature, -1), 1, generator=gen))
            else:
                nxt = int(logits.argmax())
            out.append(nxt)
            if nxt == tok.eos_id or len(out) == n_new:
                break
            r = hf(input_ids=torch.tensor([[nxt]], dtype=torch.long), past_key_values=past, use_cache=True)
            past = r.past_key_values
            logits = r.logits[0, -1]
        t_decode = time.perf_counter() - t1
        layer_decode = _tls.layer_s
    with _timings_mu:
        LAST_TIMINGS.append({"prefill_s": t_prefill, "prefill_layer_s": layer_prefill, "decode_s": t_decode,
                             "decode_layer_s": layer_decode, "prompt_tokens": len(ids), "new_tokens": len(out),
                             "n_layers": spec.n_layers})
    body = out[:-1] if out and out[-1] == tok.eos_id else out
    return SimpleNamespace(
        choices=[SimpleNamespace(message=SimpleNamespace(content=tok.decode(body)))],
        usage=SimpleNamespace(prompt_tokens=len(ids), completion_tokens=len(out)))
