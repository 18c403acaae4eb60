#!/bin/bash
# one GPU: seed fraction of the whole-table seeding of a first segment (1/12 default, 1/16, 1/24, 1/32), 8 segments
mkdir -p gpurun_out
for dv in 8 6; do timeout 200 python profiles/probe_segments.py 8 20 $dv; done > gpurun_out/probe_segments_seed_div.json; cat gpurun_out/probe_segments_seed_div.json
