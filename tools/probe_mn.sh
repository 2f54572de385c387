#!/bin/bash
# run the MN-major probe over descriptor variants; each in its own process (a trap kills only that one)
cd "$(dirname "$0")/.."
run() { echo "=== $*"; env "$@" timeout 120 python tools/probe_mn.py 2>&1 | tail -14; }
run ICL_MN_LAYOUT=1 ICL_MN_SBO=512 ICL_MN_TMASW=4
run ICL_MN_LAYOUT=1 ICL_MN_SBO=1024 ICL_MN_TMASW=4
run ICL_MN_LAYOUT=1 ICL_MN_SBO=512 ICL_MN_LBO=512 ICL_MN_TMASW=4
run ICL_MN_LAYOUT=2 ICL_MN_SBO=1024 ICL_MN_TMASW=3
run ICL_MN_LAYOUT=1 ICL_MN_SBO=512 ICL_MN_TMASW=5
