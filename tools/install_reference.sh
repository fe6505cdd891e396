#!/usr/bin/env bash
# Installs the UNMODIFIED reference package into baseline/_ref (git-ignored, NOT gpurun-ignored: it travels to the
# GPU box) so that tools/bench_triton_ref.py can time its Triton kernels on the same B200.  Run in the build
# container, where /root/reference is mounted read-only (the build writes an egg-info, hence the /tmp copy).
# Nothing in the product imports baseline/_ref.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
TMP="$(mktemp -d)"
cp -r "$SRC" "$TMP/ref"
rm -rf "$ROOT/baseline/_ref"
mkdir -p "$ROOT/baseline"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$ROOT/baseline/_ref" "$TMP/ref"
rm -rf "$TMP"
ls "$ROOT/baseline/_ref"
