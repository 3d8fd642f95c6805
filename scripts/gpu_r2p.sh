#!/bin/bash
# Round 2, visit P: store-pattern micro-benchmark (how a row-per-thread epilogue should write its tile).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 scripts/micro/store_patterns > gpurun_out/store_patterns.log 2>&1
echo "[store patterns] exit $?"; cat gpurun_out/store_patterns.log
