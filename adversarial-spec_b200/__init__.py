"""B200-native engine behind adversarial-spec's opponent-critique fan-out.

The directory name carries a hyphen (it mirrors the reference's repo name), so
import it through ``advspec_loader.load()`` at the repo root, which registers
this package as ``advspec_b200``.
"""
