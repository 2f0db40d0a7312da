"""Identity of the kernel sources a profile was taken from: bench.py refuses ncu-derived figures (profiles/traffic.json) whose
`source_hash` differs from the sources it is running."""
from __future__ import annotations

import hashlib
from pathlib import Path

_ROOT = Path(__file__).resolve().parent


def source_hash() -> str:
    h = hashlib.sha256()
    files = sorted(list((_ROOT / "csrc").glob("*.cu")) + list((_ROOT / "csrc").glob("*.cuh")) + list((_ROOT / "csrc").glob("*.h")) +
                   list((_ROOT.parent / "include").glob("*.h")))
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()[:16]
