#!/bin/bash
mkdir -p gpurun_out
NEV=60 ICL_TRACE_CS=8 python tools/trace_bptt.py 0 > gpurun_out/r2f_trace_bptt_cs8.txt 2>&1
NEV=60 ICL_TRACE_CS=4 python tools/trace_bptt.py 0 > gpurun_out/r2f_trace_bptt_cs4.txt 2>&1
head -30 gpurun_out/r2f_trace_bptt_cs8.txt
