set -x
B="python bench.py --steps 2 --warmup 3 --skip-c5 --no-cpu-baseline"
$B > gpurun_out/r02_ncu_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_bench.csv $B > gpurun_out/r02_ncu_bench.log 2>&1
python tools/run_pod_once.py c2 gram2 > gpurun_out/r02_plain_gram2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gram_streamk -s 1 -c 1 -o gpurun_out/r02_gram python tools/run_pod_once.py c2 gram2 > gpurun_out/r02_ncu_gram.log 2>&1
python tools/run_pod_once.py c2 tsqr > gpurun_out/r02_plain_tsqr.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_tsqr_c2.csv python tools/run_pod_once.py c2 tsqr > gpurun_out/r02_ncu_tsqr_l.log 2>&1
python tools/run_pod_once.py c2 tsqr > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:qr_panel -s 2 -c 2 -o gpurun_out/r02_qr_panel python tools/run_pod_once.py c2 tsqr > gpurun_out/r02_ncu_tsqr.log 2>&1
ls -la gpurun_out | tail -15
