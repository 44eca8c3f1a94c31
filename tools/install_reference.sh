#!/usr/bin/env bash
# Installs the UNMODIFIED reference package (pinton-lab/fullwave25 v1.0.16) into baseline/_ref
# (git-ignored; travels to the GPU box with the gpurun snapshot) for bench.py's reference arm and
# for tools/make_ref_golden.py.
#
# The sanctioned command
#   python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse \
#          --target baseline/_ref /root/reference
# fails in this image because the reference's build backend (hatchling) is not installed and not in
# the wheelhouse.  hatchling only zips the `fullwave` package tree into a wheel, so this script
# installs from a copy under /tmp whose pyproject.toml names setuptools as the packager instead;
# every file of the `fullwave` package is installed byte-for-byte.  To keep the snapshot small only
# the engine binaries this hardware can run (sm_100, CUDA 12.9; 6 of the 186 files) are staged.
set -euo pipefail
REPO="$(cd "$(dirname "$0")/.." && pwd)"
SRC=/root/reference
TMP=/tmp/fw25_ref_src
rm -rf "$TMP" "$REPO/baseline/_ref"
mkdir -p "$TMP"
# package tree without the 184 MB of per-architecture binaries ...
tar -C "$SRC" --exclude='fullwave/solver/bins/gpu' --exclude='fullwave/solver/bins/exponential_attenuation' \
    -cf - fullwave README.md LICENSE | tar -C "$TMP" -xf -
# ... plus the sm_100 / CUDA 12.9 executables
( cd "$SRC" && find fullwave/solver/bins -type f -name '*sm_100_cuda129' -print0 | \
    tar --null -T - -cf - ) | tar -C "$TMP" -xf -
cat > "$TMP/pyproject.toml" <<'TOML'
[project]
name = "fullwave25"
version = "1.0.16"
requires-python = ">=3.10"
[build-system]
requires = ["setuptools"]
build-backend = "setuptools.build_meta"
[tool.setuptools.packages.find]
include = ["fullwave*"]
[tool.setuptools.package-data]
"*" = ["**/*"]
TOML
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$REPO/baseline/_ref" "$TMP" 2>&1 | tail -3
# the shipped example scripts (BASELINE.json configs 1-4) are not part of the wheel: staged next to the package so that
# bench.py --config examples can run them verbatim on the GPU box (baseline/_ref is git-ignored; nothing is copied
# into the repository's history)
rm -rf "$REPO/baseline/_ref/examples"
tar -C "$SRC" --exclude='*.ipynb' -cf - examples | tar -C "$REPO/baseline/_ref" -xf -
# the relaxation-parameter database is missing from the reference checkout (.MISSING_LARGE_BLOBS): a stand-in with the
# same schema goes where fullwave.Medium looks for it -- into baseline/_ref ONLY (fullwave25_b200/lut_standin.py)
( cd "$REPO" && python -c "from fullwave25_b200 import lut_standin; print('stand-in LUT:', lut_standin.install_into('baseline/_ref'))" )
# pip drops the executable bit of package data on some versions; the launcher needs it
find "$REPO/baseline/_ref/fullwave/solver/bins" -type f -name 'fullwave2_*' -exec chmod +x {} +
echo "installed: $(find "$REPO/baseline/_ref" -type f | wc -l) files, $(du -sh "$REPO/baseline/_ref" | cut -f1)"
