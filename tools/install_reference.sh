#!/bin/bash
# Install the UNMODIFIED reference into baseline/_ref (git-ignored; travels to the GPU box with gpurun) and put a
# copy of its example scripts beside it so they can be run there (they write figures / CSV next to themselves).
# Used by bench.py --impl reference / cpu_baseline (oracle/ref_loader.py) and tests/test_examples_gpu.py.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${OPENGODDARD_REF:-/root/reference}"
[ -f "$SRC/OpenGoddard/optimize.py" ] || { echo "no reference tree at $SRC"; exit 0; }
TMP="$(mktemp -d)"
cp -r "$SRC" "$TMP/ref"            # /root/reference is read-only: build from a copy
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$ROOT/baseline/_ref" --upgrade "$TMP/ref"
mkdir -p "$ROOT/baseline/_ref/examples"
cp "$SRC"/examples/*.py "$ROOT/baseline/_ref/examples/"
# data tables the scripts read (example 11) and the (empty) output directories they write figures into
for d in "$SRC"/examples/*/; do
    n="$(basename "$d")"; mkdir -p "$ROOT/baseline/_ref/examples/$n"
    cp "$d"*.csv "$ROOT/baseline/_ref/examples/$n/" 2>/dev/null || true
done
rm -rf "$TMP"
echo "reference installed under $ROOT/baseline/_ref"
