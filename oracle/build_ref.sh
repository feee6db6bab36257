#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the unmodified PhyloStar/CyBayes reference
# (Cython extension modules + byte-compiled drivers) from the sources where they
# lie under /root/reference into oracle/_ref/ (git-ignored, travels with gpurun).
# Nothing under oracle/ may be imported by the product path (cybayes_b200/).
#
# The reference builds in place (setup.py build_ext --inplace, setup.py:7-8), and
# /root/reference is read-only, so the build happens in a throw-away temp dir and
# only the compiled artefacts (.so, byte-compiled .code) are kept.  No reference source file is
# copied into the repository.
set -euo pipefail
REF=${1:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "build_ref: $REF not present (GPU box?) - keeping prebuilt $OUT" >&2
  exit 0
fi
TMP="$(mktemp -d /tmp/cybayes_ref_build.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$REF"/. "$TMP"/
chmod -R u+w "$TMP"
( cd "$TMP" && python3 setup.py build_ext --inplace > "$TMP/build.log" 2>&1 ) || { tail -30 "$TMP/build.log"; exit 1; }
mkdir -p "$OUT"
rm -f "$OUT"/*.so "$OUT"/*.pyc "$OUT"/*.code
for m in ML ML_gamma mcmc mcmc_gamma utils config; do
  cp "$TMP"/$m.*.so "$OUT"/
done
# byte-compile the two live driver scripts (compiled artefact, not source)
python3 - "$TMP" "$OUT" <<'PY'
import py_compile, sys
tmp, out = sys.argv[1:3]
for drv in ("mat_mcmc_gamma", "mat_mcmc"):
    py_compile.compile(f"{tmp}/{drv}.py", cfile=f"{out}/{drv}.code", doraise=True)
PY
echo "build_ref: wrote $(ls "$OUT" | tr '\n' ' ')"
