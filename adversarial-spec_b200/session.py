"""Session + checkpoint side effects of a critique round.

``run_critique`` triggers these (skills/adversarial-spec/scripts/debate.py:855-878),
so the on-disk formats are kept: ``~/.config/adversarial-spec/sessions/<id>.json``
with the SessionState fields (session.py:16-40) and
``$CWD/.adversarial-spec-checkpoints/[<id>-]round-N.md`` (session.py:74-82).
No model or KV state is persisted.
"""

from __future__ import annotations

import json
import sys
from dataclasses import asdict, dataclass, field
from datetime import datetime
from pathlib import Path
from typing import Optional

SESSIONS_DIR = Path.home() / ".config" / "adversarial-spec" / "sessions"
CHECKPOINTS_DIR = Path.cwd() / ".adversarial-spec-checkpoints"


def _inside(path: Path, root: Path, what: str) -> Path:
    if not path.resolve().is_relative_to(root.resolve()):
        raise ValueError(f"Invalid session ID: {what}")
    return path


@dataclass
class SessionState:
    session_id: str
    spec: str
    round: int
    doc_type: str
    models: list
    focus: Optional[str] = None
    persona: Optional[str] = None
    preserve_intent: bool = False
    created_at: str = ""
    updated_at: str = ""
    history: list = field(default_factory=list)

    def save(self) -> None:
        SESSIONS_DIR.mkdir(parents=True, exist_ok=True)
        self.updated_at = datetime.now().isoformat()
        _inside(SESSIONS_DIR / f"{self.session_id}.json", SESSIONS_DIR, self.session_id).write_text(
            json.dumps(asdict(self), indent=2))

    @classmethod
    def load(cls, session_id: str) -> "SessionState":
        path = _inside(SESSIONS_DIR / f"{session_id}.json", SESSIONS_DIR, session_id)
        if not path.exists():
            raise FileNotFoundError(f"Session '{session_id}' not found")
        return cls(**json.loads(path.read_text()))


def save_checkpoint(spec: str, round_num: int, session_id: Optional[str] = None) -> None:
    CHECKPOINTS_DIR.mkdir(parents=True, exist_ok=True)
    prefix = f"{session_id}-" if session_id else ""
    path = _inside(CHECKPOINTS_DIR / f"{prefix}round-{round_num}.md", CHECKPOINTS_DIR, str(session_id))
    path.write_text(spec)
    print(f"Checkpoint saved: {path}", file=sys.stderr)
