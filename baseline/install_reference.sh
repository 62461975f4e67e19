#!/usr/bin/env bash
# Install the UNMODIFIED reference (papalotis/ft-fsd-path-planning) into baseline/_ref for `bench.py --impl reference`
# and the `cpu_baseline.reference_numba` leg.  baseline/_ref is git-ignored but NOT gpurun-ignored: it travels to the
# GPU box with the repository snapshot (the box has numpy / scipy / numba / scikit-learn, same image, no network).
#
#   bash baseline/install_reference.sh [/root/reference]
#
# Outcome recorded in DESIGN.md section 8:
#   1. the contract's command (pip install --no-index --target baseline/_ref <copy of the reference>) succeeds but
#      installs only the top-level package: the reference's pyproject.toml says `packages = ["fsd_path_planning"]`,
#      which leaves out every sub-package (sorting_cones/, cone_matching/, calculate_path/, utils/ ...);
#   2. so the same pip install is repeated from a /tmp copy whose [tool.setuptools] table lists the sub-packages
#      (packaging metadata only -- no .py file of the reference is touched), which yields the complete package;
#   3. `icecream` (an unused debugging import, functional_cone_matching.py:15) is neither installed nor in the
#      wheelhouse: a one-line stub module is placed next to the package.
set -euo pipefail
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
DST="$HERE/_ref"
TMP="$(mktemp -d /tmp/fsd_ref_XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
[ -d "$SRC/fsd_path_planning" ] || { echo "reference not found under $SRC" >&2; exit 1; }
cp -r "$SRC" "$TMP/src"
rm -rf "$DST"
PIP="python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --target $DST"
$PIP "$TMP/src" >"$TMP/pip1.log" 2>&1 || { cat "$TMP/pip1.log" >&2; exit 1; }
if [ ! -d "$DST/fsd_path_planning/sorting_cones" ]; then
  echo "stock install is incomplete (sub-packages missing from the reference's pyproject.toml): re-installing with package discovery"
  python - "$TMP/src/pyproject.toml" <<'EOF'
import re, sys
p = sys.argv[1]
s = open(p).read()
s = s.replace('packages = ["fsd_path_planning"]', '')
s += '\n[tool.setuptools.packages.find]\ninclude = ["fsd_path_planning*"]\n\n[tool.setuptools.package-data]\n"*" = ["*.json", "*.npy", "*.txt"]\n'
open(p, "w").write(s)
EOF
  rm -rf "$DST" "$TMP/src/build" "$TMP/src"/*.egg-info
  $PIP "$TMP/src" >"$TMP/pip2.log" 2>&1 || { cat "$TMP/pip2.log" >&2; exit 1; }
fi
[ -d "$DST/fsd_path_planning/sorting_cones" ] || { echo "install failed: sub-packages still missing" >&2; exit 1; }
cat >"$DST/icecream.py" <<'EOF'
"""Stub for the reference's unused `from icecream import ic` (not installed, not in the wheelhouse)."""


def ic(*args, **kwargs):
    return args[0] if args else None
EOF
# every .py file of the installed package must be byte-identical to the reference's
( cd "$SRC" && find fsd_path_planning -name '*.py' | sort | while read -r f; do
    [ -f "$DST/$f" ] && cmp -s "$f" "$DST/$f" || { echo "MISMATCH or missing: $f" >&2; exit 1; }
  done )
find "$DST" -name '__pycache__' -prune -exec rm -rf {} +
echo "reference installed into $DST ($(find "$DST/fsd_path_planning" -name '*.py' | wc -l) .py files, unmodified)"
