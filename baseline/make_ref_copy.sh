#!/bin/sh
# Unmodified copy of the reference for tools/reference_gpu_compare.py (see README.md here).
set -e
here="$(cd "$(dirname "$0")" && pwd)"
rm -rf "$here/_ref"
mkdir -p "$here/_ref"
cp -r /root/reference/modules /root/reference/run_ggs.py /root/reference/run_sags.py "$here/_ref/"
find "$here/_ref" -name __pycache__ -prune -exec rm -rf {} +
chmod -R u+w "$here/_ref"
echo "copied to $here/_ref"
