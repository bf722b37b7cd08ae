#!/bin/bash
# GPU call: strip / half2 producers of conv_tc6, tcgen05 attention as the default, long enough graph-replay benches to separate
# a transient from a steady-state difference.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -m gpu -s -k "variants_agree or per_engine or taps or full_size_forward or file_service or range" > gpurun_out/gpu_tests5.log 2>&1
timeout 300 python tools/check_attention.py > gpurun_out/check_attention5.log 2>&1
timeout 900 python tools/ab_forward.py tc6_lean=1 tc6_lean=2 tc6_lean=3 attn_variant=3 tc6_lean=2,fir_variant=2,outconv_variant=3,inconv_variant=2,combine_variant=1,tc1_narrow=1,gn_self=1,gnfin_variant=1 > gpurun_out/ab_lean5.log 2>&1
M=gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum
for L in 2 3; do timeout 300 ncu --clock-control none -k regex:conv_tc6 -s 54 -c 4 --metrics $M --csv --log-file gpurun_out/lean$L.csv python tools/profile_forward.py --batch 16 --evals 2 --opt tc6_lean=$L > gpurun_out/lean$L.log 2>&1; done
for V in default tc6_lean=1 tc6_lean=2 tc6_lean=3; do
  O=""; [ "$V" != "default" ] && O="--opt $V"
  timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-roofline $O > gpurun_out/bench5_$V.json 2> gpurun_out/bench5_$V.err
done
timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-roofline --opt tc6_lean=2 --opt fir_variant=2 --opt outconv_variant=3 --opt inconv_variant=2 --opt combine_variant=1 --opt tc1_narrow=1 --opt gn_self=1 --opt gnfin_variant=1 > gpurun_out/bench5_lean2_cands.json 2> gpurun_out/bench5_lean2_cands.err
tail -4 gpurun_out/gpu_tests5.log; cat gpurun_out/check_attention5.log gpurun_out/ab_lean5.log; for f in gpurun_out/bench5_*.json; do echo $f; cut -c1-120 $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'])"; done
