#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; cfg=$2; shift 2; env "$@" timeout 300 python tools/probes/gpu_stage_probe.py $cfg 2>&1 | tail -1 | cut -c1-125 | sed "s/^/[$tag] /"; }
{
cfg="config3_gp 10000 8192"
run default "$cfg" VELA_B200_JIT_DUMP=gpurun_out/k2_default.cubin
run nodelta "$cfg" VELA_B200_DELTA=0 VELA_B200_JIT_DUMP=gpurun_out/k2_nodelta.cubin
run nospec "$cfg" VELA_B200_JIT_SPECULATE=0
run nospec_nodelta "$cfg" VELA_B200_JIT_SPECULATE=0 VELA_B200_DELTA=0
run default_again "$cfg" X=1
} > gpurun_out/k2_stages.log 2>&1
cat gpurun_out/k2_stages.log
python - <<'PY'
import sys
sys.path.insert(0,'.')
from __graft_entry__ import load_package
vb=load_package()
ds=vb.synth.make_dataset("config3_gp", lambda m,t: vb.Pulsar(m,t), ntoa=2000)
psr=vb.Pulsar(ds.model, ds.toas)
print(psr.jit_log()[:1500])
PY
