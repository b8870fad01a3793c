#!/bin/bash
# compute-sanitizer over the small-shape cases of tools/sanitize_cases.py.  ONE tool per gpurun call
# (/opt/skills/guides/B200_PROFILING.md): usage  tools/sanitize.sh memcheck|racecheck|synccheck|initcheck
#   gpurun --timeout 1500 -- bash tools/sanitize.sh memcheck
# The plain run goes first (a faulting program must not be put under the tool); logs land in gpurun_out/ and the
# summaries worth keeping are copied to profiles/ by hand.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"; mkdir -p gpurun_out
TOOL="${1:-memcheck}"
timeout 600 python tools/sanitize_cases.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1400 compute-sanitizer --tool "$TOOL" --print-limit 20 --error-exitcode 7 \
    python tools/sanitize_cases.py > "gpurun_out/sanitize_${TOOL}.log" 2>&1
echo "compute-sanitizer --tool $TOOL rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|err=|sanitize cases ok|Error|hazard" "gpurun_out/sanitize_${TOOL}.log" | tail -40
