#!/bin/bash
# TEST / BASELINE INFRASTRUCTURE ONLY.  Recipe for oracle/_ref: the UNMODIFIED reference modules of the hot path, copied from
# where they lie (default /root/reference) so that `bench.py --impl reference` can time the reference's own code
# (cpu_baseline.kind = "reference") on the GPU box, where /root/reference does not exist.  oracle/_ref/ is git-ignored (no
# reference source enters the history) but not gpurun-ignored (it travels with the snapshot like the built .so).
# Files: flow/*.py (flow/flow.py, mobiusflow.py, condition.py, affineflow.py, squeezetrans.py, rottrans.py), utils/sd.py,
# utils/fisher.py, settings/*.yml.  pytorch3d / healpy / nflows / tkinter are absent from this image: oracle/stubs.py provides
# them (public formulas), exactly as for the golden vectors (oracle/make_golden.py).
set -e
SRC="${1:-/root/reference}"
DST="$(cd "$(dirname "$0")" && pwd)/_ref"
if [ ! -d "$SRC/flow" ]; then
  echo "make_ref: no reference at $SRC (nothing to do; bench.py --impl reference falls back to the oracle port)"
  exit 0
fi
mkdir -p "$DST/flow" "$DST/utils" "$DST/settings"
cp "$SRC"/flow/*.py "$DST/flow/"
cp "$SRC/utils/sd.py" "$SRC/utils/fisher.py" "$DST/utils/"
cp "$SRC"/settings/*.yml "$DST/settings/"
( cd "$SRC" && sha256sum flow/*.py utils/sd.py utils/fisher.py ) > "$DST/SHA256SUMS"
echo "make_ref: copied $(ls "$DST"/flow/*.py | wc -l) flow modules + utils/sd.py, utils/fisher.py, settings into $DST"
