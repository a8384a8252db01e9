#!/bin/bash
# Builds a writable overlay of the (read-only, python) reference package under /tmp so that it can be imported in
# THIS container to generate golden fixtures. Nothing from the reference is copied into the repository.
set -e
DST=${1:-/tmp/pttools_ref}
mkdir -p "$DST"
if [ ! -d "$DST/pttools" ]; then cp -r /root/reference/pttools "$DST/pttools"; chmod -R u+w "$DST/pttools"; fi
echo "export PYTHONPATH=$DST:$(cd "$(dirname "$0")" && pwd)/stubs:\$PYTHONPATH NUMBA_THREADING_LAYER=workqueue"
