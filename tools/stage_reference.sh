#!/bin/bash
# Stages an UNMODIFIED copy of the reference files that tests/test_ref_script.py executes through the shim under baseline/_ref/
# (git-ignored -- never part of the history -- but not gpurun-ignored, so it travels to the GPU box, where /root/reference does
# not exist).  Run in the build container before a gpurun call that includes tests/test_ref_script.py.
set -e
cd "$(dirname "$0")/.."
dst=baseline/_ref/ImageCaptionLearn_py
mkdir -p $dst/nn_utils $dst/utils
cp /root/reference/icl_core_lstm.py /root/reference/icl_relation_lstm.py /root/reference/icl_affinity_lstm.py $dst/
cp /root/reference/nn_utils/*.py $dst/nn_utils/
cp /root/reference/utils/*.py $dst/utils/
echo "staged $(find $dst -name '*.py' | wc -l) files under $dst"
