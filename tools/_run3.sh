for v in r1 r2; do
  export NPBNN_B200_LIB=$PWD/npbnn_b200/libnpbnn_b200_$v.so
  echo "== $v" >> gpurun_out/u5_ring.log
  python tools/fwd_bench.py 1000000 32 1 2>&1 | tail -30 >> gpurun_out/u5_ring.log
done
