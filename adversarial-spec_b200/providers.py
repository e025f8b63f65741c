"""Provider registry for the `providers` command and the credential pre-flight.

Mirrors the observable behaviour of the reference's registry for the critique
path (skills/adversarial-spec/scripts/providers.py: list_providers :247-333,
validate_model_credentials :418-486, MODEL_COSTS :18-45): same row format, same
[set]/[not set] logic, same prefix -> key rules, unknown prefixes pass and are
priced at the $5/$15 default.  Added: the ``b200/`` (alias ``local/``) provider
served by this repo's engine — free, no key.  Bedrock and profiles are out of
scope (config CRUD) and report "not configured".
"""

from __future__ import annotations

import os
import shutil
from typing import Optional

from .model_spec import LOCAL_PREFIXES, REGISTRY, is_local_model

_PRICES = [
    ("gpt-4o", 2.50, 10.00), ("gpt-4-turbo", 10.00, 30.00), ("gpt-4", 30.00, 60.00),
    ("gpt-3.5-turbo", 0.50, 1.50), ("o1", 15.00, 60.00), ("o1-mini", 3.00, 12.00),
    ("claude-sonnet-4-20250514", 3.00, 15.00), ("claude-opus-4-20250514", 15.00, 75.00),
    ("gemini/gemini-2.0-flash", 0.075, 0.30), ("gemini/gemini-pro", 0.50, 1.50),
    ("xai/grok-3", 3.00, 15.00), ("xai/grok-beta", 5.00, 15.00), ("mistral/mistral-large", 2.00, 6.00),
    ("groq/llama-3.3-70b-versatile", 0.59, 0.79), ("deepseek/deepseek-chat", 0.14, 0.28),
    ("zhipu/glm-4", 1.40, 1.40), ("zhipu/glm-4-plus", 7.00, 7.00),
    ("codex/gpt-5.2-codex", 0.0, 0.0), ("codex/gpt-5.1-codex-max", 0.0, 0.0),
    ("codex/gpt-5.1-codex-mini", 0.0, 0.0), ("gemini-cli/gemini-3-pro-preview", 0.0, 0.0),
    ("gemini-cli/gemini-3-flash-preview", 0.0, 0.0),
]
MODEL_COSTS = {name: {"input": i, "output": o} for name, i, o in _PRICES}
for _p in LOCAL_PREFIXES:  # local opponents cost nothing per token
    for _n in REGISTRY:
        MODEL_COSTS[_p + _n] = {"input": 0.0, "output": 0.0}
DEFAULT_COST = {"input": 5.00, "output": 15.00}

CODEX_AVAILABLE = shutil.which("codex") is not None
GEMINI_CLI_AVAILABLE = shutil.which("gemini") is not None
DEFAULT_CODEX_REASONING = "xhigh"

# (display name, env var, example models, default model)
_API_PROVIDERS = [
    ("OpenAI", "OPENAI_API_KEY", "gpt-4o, gpt-4-turbo, o1", "gpt-4o"),
    ("Anthropic", "ANTHROPIC_API_KEY", "claude-sonnet-4-20250514, claude-opus-4-20250514",
     "claude-sonnet-4-20250514"),
    ("Google", "GEMINI_API_KEY", "gemini/gemini-2.0-flash, gemini/gemini-pro", "gemini/gemini-2.0-flash"),
    ("xAI", "XAI_API_KEY", "xai/grok-3, xai/grok-beta", "xai/grok-3"),
    ("Mistral", "MISTRAL_API_KEY", "mistral/mistral-large, mistral/codestral", "mistral/mistral-large"),
    ("Groq", "GROQ_API_KEY", "groq/llama-3.3-70b-versatile", "groq/llama-3.3-70b-versatile"),
    ("Together", "TOGETHER_API_KEY", "together_ai/meta-llama/Llama-3-70b", None),
    ("OpenRouter", "OPENROUTER_API_KEY", "openrouter/openai/gpt-4o, openrouter/anthropic/claude-3.5-sonnet",
     None),
    ("Deepseek", "DEEPSEEK_API_KEY", "deepseek/deepseek-chat", "deepseek/deepseek-chat"),
    ("Zhipu", "ZHIPUAI_API_KEY", "zhipu/glm-4, zhipu/glm-4-plus", "zhipu/glm-4"),
]
_KEY_FOR_PREFIX = [
    ("gpt-", "OPENAI_API_KEY"), ("o1", "OPENAI_API_KEY"), ("claude-", "ANTHROPIC_API_KEY"),
    ("gemini/", "GEMINI_API_KEY"), ("xai/", "XAI_API_KEY"), ("mistral/", "MISTRAL_API_KEY"),
    ("groq/", "GROQ_API_KEY"), ("deepseek/", "DEEPSEEK_API_KEY"), ("zhipu/", "ZHIPUAI_API_KEY"),
]


def get_bedrock_config() -> dict:
    return {}


def list_providers() -> None:
    print("Supported providers:\n")
    for name, key, models, _ in _API_PROVIDERS:
        status = "[set]" if os.environ.get(key) else "[not set]"
        print(f"  {name:12} {key:24} {status}")
        print(f"             Example models: {models}")
        print()
    codex = "[installed]" if CODEX_AVAILABLE else "[not installed]"
    print(f"  {'Codex CLI':12} {'(ChatGPT subscription)':24} {codex}")
    print("             Example models: codex/gpt-5.2-codex, codex/gpt-5.1-codex-max")
    print("             Reasoning: --codex-reasoning (minimal, low, medium, high, xhigh)")
    print("             Install: npm install -g @openai/codex && codex login")
    print()
    gem = "[installed]" if GEMINI_CLI_AVAILABLE else "[not installed]"
    print(f"  {'Gemini CLI':12} {'(Google account)':24} {gem}")
    print("             Example models: gemini-cli/gemini-3-pro-preview, gemini-cli/gemini-3-flash-preview")
    print("             Install: npm install -g @google/gemini-cli && gemini auth")
    print()
    print("AWS Bedrock:\n")
    print("  Not configured. Enable with: python3 debate.py bedrock enable --region us-east-1")
    print()
    print("Local B200 engine:\n")
    print(f"  {'B200 local':12} {'(no key; CUDA sm_100a)':24} {local_engine_status()}")
    print("             Example models: " + ", ".join("b200/" + n for n in list(REGISTRY)[:6]))
    print("             Same-model opponents share one prefill and one weight stream per GPU")
    print()


def local_engine_status() -> str:
    from . import engine

    return "[built]" if engine.LIB_PATH.exists() else "[not built]"


def get_available_providers() -> list[tuple[str, str, str]]:
    return [(n, k, d) for n, k, _, d in _API_PROVIDERS if d and os.environ.get(k)]


def get_default_model() -> Optional[str]:
    avail = get_available_providers()
    return avail[0][2] if avail else None


def validate_model_credentials(models: list[str]) -> tuple[list[str], list[str]]:
    valid, invalid = [], []
    for m in models:
        if is_local_model(m):
            (valid if local_model_known(m) else invalid).append(m)
        elif m.startswith("codex/"):
            (valid if CODEX_AVAILABLE else invalid).append(m)
        elif m.startswith("gemini-cli/"):
            (valid if GEMINI_CLI_AVAILABLE else invalid).append(m)
        else:
            key = next((k for p, k in _KEY_FOR_PREFIX if m.startswith(p)), None)
            (valid if key is None or os.environ.get(key) else invalid).append(m)
    return valid, invalid


def local_model_known(model: str) -> bool:
    from .model_spec import resolve

    try:
        resolve(model)
        return True
    except KeyError:
        return False


def required_key_hint(model: str) -> str:
    if is_local_model(model):
        return "unknown local model; known: " + ", ".join("b200/" + n for n in REGISTRY)
    if model.startswith("codex/"):
        return "requires Codex CLI: npm install -g @openai/codex && codex login"
    key = next((k for p, k in _KEY_FOR_PREFIX if model.startswith(p)), None)
    return f"requires {key}" if key else "unknown provider"
