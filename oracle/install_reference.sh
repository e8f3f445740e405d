#!/bin/sh
# TEST / BENCH INFRASTRUCTURE — installs the UNMODIFIED reference package into baseline/_ref (git-ignored; it travels to
# the GPU box with the gpurun snapshot) so that `bench.py --impl reference` and the cpu_baseline leg can time the
# reference's own module there.  /root/reference is read-only and its build writes an egg-info, hence the copy.
#   --no-deps                  pystow / indra / pybel / mlflow ... are not in the offline wheelhouse (the hot path does
#                              not need them: oracle/ref_shim.py stubs the three modules its import chain touches)
#   --ignore-requires-python   setup.cfg pins python <3.9; the code of the path runs unchanged on 3.12
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
TMP="$(mktemp -d)"
cp -r /root/reference "$TMP/ref"
python -m pip install --no-index --no-build-isolation --no-deps --ignore-requires-python \
    --find-links /opt/wheelhouse --target "$ROOT/baseline/_ref" "$TMP/ref"
rm -rf "$TMP"
