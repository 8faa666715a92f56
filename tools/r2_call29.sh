#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_run.py > gpurun_out/c29_sanitize_$tool.txt 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done|Traceback|Error" gpurun_out/c29_sanitize_$tool.txt | tail -4
done
