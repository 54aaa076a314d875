#!/usr/bin/env bash
# Install the UNMODIFIED reference (PABannier/nanograd, /root/reference) into the git-ignored,
# gpurun-shipped directory oracle/_ref/ so that it can travel to the GPU box, where it serves as
#   * the timed CPU arm of `bench.py --impl reference` (cpu_baseline.kind = "reference"), and
#   * the host package of tests/test_integration_reference.py (the reference's own Tensor /
#     Module / Optimizer bound to the CUDA backend by nanograd_b200.integrate).
# This is a pip install of the reference's own setup.py (from a scratch copy, because
# /root/reference is read-only and setuptools writes build/ and egg-info next to setup.py);
# --no-deps because matplotlib / graphviz / pyopencl / mako are not in the offline wheelhouse
# (the arithmetic only needs NumPy; tests/_shim carries a 4-line graphviz stand-in).
# Nothing from the reference is committed: oracle/_ref/ is listed in .gitignore.
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
REF="${1:-/root/reference}"
DEST="$ROOT/oracle/_ref"
if [ ! -f "$REF/setup.py" ]; then
  echo "reference not present at $REF: keeping $DEST as it is" >&2
  exit 0
fi
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$TMP/src"
# sources only: the checked-in build/ and dist/ trees of the reference are stale copies
(cd "$REF" && tar cf - --exclude=build --exclude=dist --exclude='*.egg-info' --exclude=notebooks --exclude=docs .) | (cd "$TMP/src" && tar xf -)
rm -rf "$DEST"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --no-compile \
  --find-links /opt/wheelhouse --target "$DEST" "$TMP/src"
test -f "$DEST/nanograd/tensor.py"
echo "installed $(ls -d "$DEST"/nanograd-*.dist-info 2>/dev/null | xargs -n1 basename) into $DEST"
