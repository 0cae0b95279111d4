set -x
python tools/run_slab_once.py 250000 4096 tsqr_full > gpurun_out/r02b_plain_slab.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:qr_tn_kernel -s 3 -c 1 -o gpurun_out/r02b_qr_tn python tools/run_slab_once.py 250000 4096 tsqr_full > gpurun_out/r02b_ncu_qr_tn.log 2>&1
python tools/bench_snapshots.py once > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ackley_grouped -s 1 -c 1 -o gpurun_out/r02b_ackley_grouped python tools/bench_snapshots.py once > gpurun_out/r02b_ncu_ackley.log 2>&1
python tools/run_pod_once.py c2 gram2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_nn_kernel -s 2 -c 1 -o gpurun_out/r02b_rotate python tools/run_pod_once.py c2 gram2 > gpurun_out/r02b_ncu_rotate.log 2>&1
ls -la gpurun_out | tail -8
