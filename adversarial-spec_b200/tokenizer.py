"""Deterministic synthetic tokenizer + spec generator.

No vocabulary or merges file exists offline, and random-init
opponents emit ids over the whole vocabulary, so the tokenizer is defined for
ANY vocab size:  0 = <bos>, 1 = <eos>, 2..257 = raw bytes, then a built-in
lexicon of word pieces (each with and without a leading space); ids past the
lexicon wrap around it so every id decodes to text.  encode(decode(ids)) is
not the identity for wrapped ids; decode(encode(text)) is exact.
"""

from __future__ import annotations

import re

import numpy as np

BOS, EOS, BYTE0 = 0, 1, 2

_WORDS = """
the a an of to and in for with on by is are be as at from that this it or not must should may can will
system service user users client server request response api endpoint data database table index cache
queue worker job task event message stream batch record schema field key value token session account
auth login password secret permission role access policy rate limit timeout retry backoff error failure
latency throughput availability durability consistency partition replica shard leader follower quorum
deploy release rollback migration version config flag metric log trace alert dashboard monitor test
unit integration load security privacy encryption compliance audit backup restore storage network
requirement goal scope constraint assumption risk dependency milestone owner stakeholder feature story
acceptance criteria priority design architecture component module interface contract protocol format
json http https grpc rest sql nosql kafka redis postgres s3 cdn dns tls jwt oauth sso
create read update delete list get set put post patch validate verify ensure provide support handle
process store send receive return fail succeed reject accept allow deny block expire rotate refresh
when if then else each every all any no only also however therefore because within across between
one two three four five ten hundred thousand million percent second seconds minute minutes hour day
p50 p95 p99 ms kb mb gb qps slo sla mvp v1 v2 id uuid url uri ip
critical high medium low new old current next previous final initial default optional required
section overview background summary details notes open questions decision alternatives tradeoffs
""".split()

_PUNCT = [".", ",", ":", ";", "-", "(", ")", "[", "]", "/", "#", "*", "`", "'", '"', "\n", "\n\n", "  ",
          "[AGREE]", "[SPEC]", "[/SPEC]", "##", "1.", "2.", "3."]


def _build_lexicon() -> list[str]:
    seen, out = set(), []
    for w in _WORDS:
        for piece in (" " + w, w, " " + w.capitalize()):
            if piece not in seen:
                seen.add(piece)
                out.append(piece)
    for p in _PUNCT:
        if p not in seen:
            seen.add(p)
            out.append(p)
    return out


LEXICON = _build_lexicon()
LEX0 = BYTE0 + 256
_PIECE_RE = re.compile(r"\[/?[A-Z]+\]| ?[A-Za-z0-9]+|\n\n|\n|  |[^\sA-Za-z0-9]| ")


class SyntheticTokenizer:
    def __init__(self, vocab_size: int):
        if vocab_size < LEX0 + 16:
            raise ValueError(f"vocab_size {vocab_size} too small for the byte range")
        self.vocab_size = vocab_size
        self.n_lex = min(len(LEXICON), vocab_size - LEX0)
        self.piece_to_id = {p: LEX0 + i for i, p in enumerate(LEXICON[: self.n_lex])}
        self.eos_id = EOS

    def encode(self, text: str, bos: bool = False) -> list[int]:
        ids = [BOS] if bos else []
        for piece in _PIECE_RE.findall(text):
            tid = self.piece_to_id.get(piece)
            if tid is not None:
                ids.append(tid)
            else:
                ids.extend(BYTE0 + b for b in piece.encode("utf-8"))
        return ids

    def decode(self, ids) -> str:
        out = bytearray()
        for t in ids:
            t = int(t)
            if t == BOS or t == EOS or t < 0:
                continue
            if t < LEX0:
                out.append(t - BYTE0)
            else:
                out.extend(LEXICON[(t - LEX0) % self.n_lex].encode("utf-8"))
        return out.decode("utf-8", errors="replace")

    def count(self, text: str) -> int:
        return len(self.encode(text))


def render_chat(system_prompt: str, user_message: str) -> str:
    """The one chat template every local opponent uses (system, then user, then the turn marker)."""
    return f"[SYSTEM]\n{system_prompt}\n[USER]\n{user_message}\n[ASSISTANT]\n"


def generate_spec(tok: SyntheticTokenizer, n_tokens: int, seed: int, title: str = "Synthetic Spec") -> str:
    """English-like spec text of EXACTLY n_tokens tokens under `tok` (seeded)."""
    rng = np.random.default_rng(seed)
    words = [w for w in _WORDS if (" " + w) in tok.piece_to_id]
    parts: list[str] = [f"# {title}\n\n"]
    n = tok.count(parts[0])
    sec = 0
    while n < n_tokens:
        if rng.random() < 0.04:
            sec += 1
            piece = f"\n\n## Section {sec}\n\n"
        else:
            k = int(rng.integers(6, 18))
            ws = rng.choice(words, size=k)
            piece = " " + " ".join(ws) + "."
            piece = piece[:1] + piece[1:2].upper() + piece[2:]
        c = tok.count(piece)
        if n + c > n_tokens:
            break
        parts.append(piece)
        n += c
    text = "".join(parts)
    for _ in range(4):
        # the reference strips stdin (debate.py:778): keep the text strip-stable at exactly n tokens
        text = text.strip()
        c = tok.count(text)
        if c == n_tokens:
            break
        if c > n_tokens:
            text = tok.decode(tok.encode(text)[:n_tokens])
        else:
            text += "".join(" " + str(w) for w in rng.choice(words, size=n_tokens - c))
    assert tok.count(text) == n_tokens and text == text.strip()
    return text
