./tools/fp64_mix > gpurun_out/u1_mix.log 2>&1
for v in "" noact chain4 chain8 chain12 actfast; do
  if [ -z "$v" ]; then unset NPBNN_B200_LIB; else export NPBNN_B200_LIB=$PWD/npbnn_b200/libnpbnn_b200_$v.so; fi
  echo "== $v" >> gpurun_out/u1_var.log
  python tools/fwd_bench.py 1000000 32 5 >> gpurun_out/u1_var.log 2>&1
done
