#!/usr/bin/env bash
# Populate baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box) with the
# reference's own loop / data modules so that the -m gpu tests and bench.py --impl reference can drive
# the UNMODIFIED model_utils.train() / val() / test(), datasets.MultiModalX and utils.sliding_window on
# the box, where /root/reference does not exist.  The reference has no setup.py / pyproject.toml, so
# "pip install /root/reference" is impossible: the four files are taken as they lie.  Nothing under
# baseline/_ref/ is ever committed (see .gitignore) and nothing in vit-cnn_b200/ imports it.
set -euo pipefail
SRC="${VITCNN_REFERENCE_ROOT:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
DST="$HERE/baseline/_ref"
if [ ! -f "$SRC/model_utils.py" ]; then
  echo "install_ref: $SRC not present (GPU box?) - keeping whatever $DST holds" >&2
  exit 0
fi
mkdir -p "$DST"
for f in utils.py datasets.py model_utils.py losses.py; do
  install -m 0644 "$SRC/$f" "$DST/$f"
done
( cd "$SRC" && sha256sum utils.py datasets.py model_utils.py losses.py ) > "$DST/SHA256SUMS"
echo "install_ref: $(ls "$DST" | tr '\n' ' ')"
