# Same-box A/B of ONE library with and without an environment switch (first the switch set = "off" arm, then unset):
#   gpurun -- 'bash tools/ab_env.sh r03u libb200sr3_wres.so B200SR3_W_RESIDENT=0 "<shape>" ...'
# Every command runs under `timeout`: a variant that deadlocks must not hold the box.
TAG=${1:-r03x}; LIB=$2; SW=$3; shift 3
mkdir -p gpurun_out
OUT=gpurun_out/${TAG}_ab.txt
: > $OUT
export B200SR3_LIB=$PWD/3d-super-resolution-face-reconstruction_b200/b200sr3/$LIB
for rep in 1 2; do
  for name in "$@"; do
    echo -n "$SW : " >> $OUT; env $SW timeout 120 python tools/halo_bench.py 32 30 "$name" 1 >> $OUT 2>&1 || echo "FAILED rc $?" >> $OUT
    echo -n "default          : " >> $OUT; timeout 120 python tools/halo_bench.py 32 30 "$name" 1 >> $OUT 2>&1 || echo "FAILED rc $?" >> $OUT
  done
done
if grep -q FAILED $OUT; then cat $OUT; exit 1; fi
env $SW timeout 300 python tools/profile_step.py 32 128 600 > gpurun_out/${TAG}_step_off.txt 2>&1
timeout 300 python tools/profile_step.py 32 128 600 > gpurun_out/${TAG}_step_on.txt 2>&1
paste <(awk '{print $1, $2}' gpurun_out/${TAG}_step_off.txt) <(awk '{print $2}' gpurun_out/${TAG}_step_on.txt) | grep -E "downs\.[12]\.|ups\.1[678]\.|total|^conv" >> $OUT
env $SW timeout 600 python bench.py --no-cpu-baseline --no-torch-baseline > gpurun_out/${TAG}_bench_off.json 2> /dev/null
timeout 600 python bench.py --no-cpu-baseline --no-torch-baseline > gpurun_out/${TAG}_bench_on.json 2> /dev/null
python -c "
import json
for arm in ('off','on'):
    a=json.load(open('gpurun_out/${TAG}_bench_'+arm+'.json')); print(arm, 'faces/s', a['value'], 'psnr', a.get('psnr_vs_ref_db'), 'max|d|', a.get('final_max_abs'))" >> $OUT
cat $OUT
