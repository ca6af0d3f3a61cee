# Same-box A/B of two library builds: isolated layers (halo_bench), the full bench, and the per-launch step profile.
#   gpurun -- 'bash tools/ab_step_round.sh r03n libb200sr3_f2.so "<shape>" ...'
TAG=${1:-r03x}; VAR=${2:-libb200sr3_f2.so}
bash tools/ab_round.sh "$@"
LIBV=$PWD/3d-super-resolution-face-reconstruction_b200/b200sr3/$VAR
python tools/profile_step.py 32 128 600 > gpurun_out/${TAG}_step_default.txt 2>&1
B200SR3_LIB=$LIBV python tools/profile_step.py 32 128 600 > gpurun_out/${TAG}_step_variant.txt 2>&1
paste <(awk '{print $1, $2}' gpurun_out/${TAG}_step_default.txt) <(awk '{print $2}' gpurun_out/${TAG}_step_variant.txt) | head -75
