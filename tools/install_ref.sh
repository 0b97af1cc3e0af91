#!/bin/bash
# Install the UNMODIFIED reference for the reference arm of bench.py and the drop-in driver tests.
#
# The reference (XiangFeng-Wen/STF-Unet) is a pure-Python tree without setup.py / pyproject.toml, so
#   python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
# fails ("does not appear to be a Python project": recorded in DESIGN.md section 5).  Its hot path needs exactly two
# packages -- src/ (the models) and train_utils/ (criterion, train_one_epoch, evaluate) -- which this script copies
# byte for byte into baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box, where /root/reference
# does not exist).  Nothing under baseline/_ref is product source; only bench.py's reference arm / gpu_eager_baseline
# leg and tests/ import it.
set -euo pipefail
SRC=${1:-/root/reference}
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
DST="$ROOT/baseline/_ref"
if [ ! -d "$SRC/src" ] || [ ! -d "$SRC/train_utils" ]; then
  echo "install_ref: $SRC has no src/ + train_utils/ (the reference is only present in the build container)" >&2
  exit 0
fi
rm -rf "$DST"
mkdir -p "$DST"
cp -r "$SRC/src" "$DST/src"
cp -r "$SRC/train_utils" "$DST/train_utils"
find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} +
( cd "$SRC" && find src train_utils -name '*.py' -print0 | sort -z | xargs -0 sha256sum ) > "$DST/SHA256SUMS"
( cd "$DST" && sha256sum -c SHA256SUMS --quiet )
echo "install_ref: $(wc -l < "$DST/SHA256SUMS") files -> $DST (sha256 verified against $SRC)"
