#!/usr/bin/env bash
# Submit a whole suite to the cluster the current kubectl context points at:
#   tests/workloads/run_all.sh short|long
set -euo pipefail
SUITE=${1:-short}
HERE=$(cd "$(dirname "$0")" && pwd)
python3 "$HERE/workloads.py" list | awk -v s="$SUITE" '$1 == s {print $2}' |
while read -r name; do
    echo "submitting $name"
    python3 "$HERE/workloads.py" submit "$name"
done
