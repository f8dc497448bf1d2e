./tools/pipe_cost a b > gpurun_out/u4_lat.log 2>&1
for v in t0 t1 t2; do
  export NPBNN_B200_LIB=$PWD/npbnn_b200/libnpbnn_b200_$v.so
  echo "== $v" >> gpurun_out/u4_var.log
  python tools/fwd_bench.py 1000000 32 5 >> gpurun_out/u4_var.log 2>&1
  python tools/fwd_bench.py 1000000 4 10 >> gpurun_out/u4_var.log 2>&1
done
export NPBNN_B200_LIB=$PWD/npbnn_b200/libnpbnn_b200_t2.so
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "c4_shape or headline or family or forward_lik or c4_full" > gpurun_out/u4_pytest.log 2>&1
