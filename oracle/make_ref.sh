#!/usr/bin/env bash
# Test/bench infrastructure, NOT product code: puts the UNMODIFIED reference (Academich/translation-transformer) hot-path
# packages next to the oracle so that `bench.py --impl reference` and the cpu_baseline leg can time the real thing on the
# GPU box (which has no /root/reference).  Runs in the build container only; output goes to oracle/_ref/, which is
# git-ignored (it is not part of this repository's history) but travels with the gpurun snapshot like the built .so files.
#
#   oracle/_ref/src/{model,utils,decoding,data_handling}   verbatim copies (cp -p, checked with cmp below)
#   oracle/_ref/stubs/pytorch_lightning/                    import stub: the reference's packages subclass Lightning
#                                                           classes at import time; Lightning is not installed here
#   oracle/_ref/MANIFEST                                    sha256 of every copied file
set -euo pipefail
REF=${TTB_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  echo "make_ref: $REF/src not found (only the build container has the reference); keeping $OUT as it is" >&2
  exit 0
fi
rm -rf "$OUT"
mkdir -p "$OUT/src" "$OUT/stubs/pytorch_lightning/utilities"
for pkg in model utils decoding data_handling; do
  cp -rp "$REF/src/$pkg" "$OUT/src/$pkg"
done
find "$OUT/src" -name '__pycache__' -type d -prune -exec rm -rf {} +
# verbatim check + manifest
( cd "$OUT/src" && find . -type f -name '*.py' | sort | while read -r f; do
    cmp -s "$f" "$REF/src/$f" || { echo "make_ref: $f differs from the reference" >&2; exit 1; }
    sha256sum "$f"
  done ) > "$OUT/MANIFEST"
cat > "$OUT/stubs/pytorch_lightning/__init__.py" <<'PY'
"""Import stub (oracle/make_ref.sh): only the names the reference's modules touch at import time."""
import torch


class LightningModule(torch.nn.Module):
    def save_hyperparameters(self, *a, **k):
        pass


class LightningDataModule:
    pass


class Callback:
    pass


class Trainer:
    pass
PY
cat > "$OUT/stubs/pytorch_lightning/utilities/__init__.py" <<'PY'
PY
cat > "$OUT/stubs/pytorch_lightning/utilities/types.py" <<'PY'
STEP_OUTPUT = object
PY
cat > "$OUT/stubs/pytorch_lightning/callbacks.py" <<'PY'
class BasePredictionWriter:
    def __init__(self, write_interval="batch"):
        self.interval = write_interval
PY
echo "make_ref: $(wc -l < "$OUT/MANIFEST") reference files under $OUT"
