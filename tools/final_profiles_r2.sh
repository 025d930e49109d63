# ncu evidence of the final round-2 code (one GPU).  The stiff pass beside the bulk pass cannot be captured while it runs
# (ncu serialises kernels): the sweep is profiled with the stiff pass AFTER the bulk pass (auto_flags = 8), same kernels.
set -x
python tools/profile_auto.py auto 8 > gpurun_out/r2u_plain.log 2>&1 && tail -1 gpurun_out/r2u_plain.log
ncu --set full --import-source on --clock-control none -k regex:odl_sweep_kernel -s 2 -c 1 -f -o gpurun_out/r2u_sweep python tools/profile_auto.py auto 8 > gpurun_out/r2u_ncu_sweep.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:odl_sweep_bdf_kernel -s 2 -c 1 -f -o gpurun_out/r2u_tail python tools/profile_auto.py auto 8 > gpurun_out/r2u_ncu_tail.log 2>&1
CHAINS=8192 ITS=20 python tools/profile_coop.py > gpurun_out/r2u_coop_plain.log 2>&1 && tail -1 gpurun_out/r2u_coop_plain.log
ncu --set full --import-source on --clock-control none -k regex:odl_mcmc_coop_kernel -s 1 -c 1 -f -o gpurun_out/r2u_coop python tools/profile_coop.py > gpurun_out/r2u_ncu_coop.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu --no-facade --no-configs --no-cold > gpurun_out/r2u_bench_plain.log 2>&1 && tail -c 300 gpurun_out/r2u_bench_plain.log
ODL_WATCHDOG_SPINS=1000 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2u_bench_launch_list.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-facade --no-configs --no-cold > gpurun_out/r2u_ncu_bench.log 2>&1
ls -la gpurun_out/r2u*
