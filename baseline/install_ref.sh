#!/usr/bin/env bash
# Offline install of the UNMODIFIED reference (facebookresearch/SlowFast) into baseline/_ref (git-ignored; it travels to
# the GPU box with the gpurun snapshot).  Run in the build container, where /root/reference exists:
#
#     bash baseline/install_ref.sh
#
# 1. the contract's pip install (the reference's own setup.py; the source tree is read-only, so build from a /tmp copy;
#    dependency resolution cannot succeed offline -> --no-deps; the un-vendored imports are served by oracle/refshim.py);
# 2. setup.py's find_packages() ships only the `slowfast` package: the driver scripts (tools/train_net.py, test_net.py),
#    the yaml configs and ava_evaluation/ are plain files of the checkout, copied verbatim next to it.
# Nothing under baseline/_ref is ever edited, and nothing of it is tracked by git.
set -euo pipefail
REF=${SLOWFAST_REFERENCE_SRC:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
DST="$HERE/_ref"
if [ ! -d "$REF/slowfast" ]; then
  echo "reference checkout not found at $REF (only the build container has it)" >&2
  exit 1
fi
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
cp -r "$REF" "$TMP/src"
rm -rf "$DST"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
  --target "$DST" "$TMP/src"
for d in tools configs ava_evaluation; do
  cp -r "$REF/$d" "$DST/$d"
done
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
( cd "$REF" && find slowfast tools -name '*.py' -print0 | sort -z | xargs -0 sha256sum ) > "$TMP/src.sha"
( cd "$DST" && find slowfast tools -name '*.py' -print0 | sort -z | xargs -0 sha256sum ) > "$TMP/dst.sha"
cmp "$TMP/src.sha" "$TMP/dst.sha"   # byte-identical to the checkout
echo "installed: $DST ($(du -sh "$DST" | cut -f1)), $(wc -l < "$TMP/dst.sha") python files identical to $REF"
