#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/step_profile.py > gpurun_out/step_profile.md 2> gpurun_out/step_profile.err
head -45 gpurun_out/step_profile.md; tail -3 gpurun_out/step_profile.err
