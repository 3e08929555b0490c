#!/bin/bash
# Installs the UNMODIFIED reference (FBPIC, /root/reference) into the git-ignored oracle/_ref/ so that
# `bench.py --impl reference` can time FBPIC's own numba CPU path on the GPU box's host cores
# (oracle/_ref travels with gpurun; /root/reference does not exist there).  Test / measurement
# infrastructure only: nothing under fbpic_b200/ imports it.  The reference's source tree is read-only and
# its setup.py imports the package, so the install runs from a scratch copy under /tmp.
#   bash oracle/make_ref.sh        (also run by __graft_entry__.build() when /root/reference is present)
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${REFERENCE_DIR:-/root/reference}
[ -d "$REF/fbpic" ] || { echo "make_ref: $REF not present, keeping the prebuilt oracle/_ref" >&2; exit 0; }
if [ -f "$HERE/_ref/fbpic/main.py" ] && [ "$HERE/_ref/fbpic/main.py" -nt "$REF/fbpic/main.py" ]; then exit 0; fi
TMP=$(mktemp -d /tmp/fbpic_ref.XXXXXX)
cp -r "$REF/." "$TMP/"
rm -rf "$HERE/_ref"
PYTHONPATH="$HERE/ref_shim" python -m pip install --quiet --no-index --no-build-isolation --no-deps \
    --find-links /opt/wheelhouse --target "$HERE/_ref" "$TMP"
rm -rf "$TMP"
echo "installed: $(ls "$HERE/_ref")"
