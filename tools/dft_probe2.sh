#!/usr/bin/env bash
# Probe matrix for the STFT GEMM (bring-up build, libavld_bringup.so): AVLD_DBG bits, see dftf3.cu.
# Results of runs with AVLD_DBG other than 0 / 64 are garbage by design; only the kernel time and the cycle counters matter.
# Usage: gpurun -- tools/dft_probe2.sh [tag ...]     (no tags = all)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export AVLD_LIB_PATH=$PWD/amphibian_vae_latent_detector_b200/libavld_bringup.so
small="--chunks 8192 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
want=" $* "
run() {  # tag, env...
  tag=$1; shift
  if [ "$want" != "  " ] && [[ "$want" != *" $tag "* ]]; then return; fi
  env "$@" AVLD_PROF_DUMP=1 timeout 300 python bench.py $small > gpurun_out/probe_$tag.log 2> gpurun_out/probe_$tag.err
  rc=$?
  python - "$tag" "$rc" <<'PY'
import json, sys
tag, rc = sys.argv[1], sys.argv[2]
try:
    line = [l for l in open(f"gpurun_out/probe_{tag}.log") if l.startswith("{")][-1]
    d = json.loads(line)
    st = d.get("stage_ms_per_step", {})
    print(f"{tag:14s} rc {rc} dft ms/launch {d['roofline']['avg_launch_ms']:.4f} fold3 ms/step {st.get('fold3_kernel')} prep {st.get('prep_kernel')} "
          f"chunks/s {d['value']:.0f} sm_mhz {d.get('clocks', {}).get('sm_mhz')}")
except Exception as e:
    print(tag, "rc", rc, "no bench line:", e)
    print(open(f"gpurun_out/probe_{tag}.err").read()[-1500:])
prof = [l.strip() for l in open(f"gpurun_out/probe_{tag}.err") if "prof (cycles" in l]
if prof:
    print("   ", prof[-1])
PY
}
run base          AVLD_DBG=0
run base_prof     AVLD_DBG=64
run noepi         AVLD_DBG=65
run noloads       AVLD_DBG=68
run no_a          AVLD_DBG=72
run no_b          AVLD_DBG=80
run onepass       AVLD_DBG=96
