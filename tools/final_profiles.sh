# ncu evidence of the final kernels (one GPU).  The stiff pass beside the bulk pass cannot be captured while it runs
# (ncu serialises kernels): the sweep is profiled with the stiff pass AFTER the bulk pass (auto_flags = 8), same kernels.
set -x
python tools/profile_auto.py auto 8 > gpurun_out/r2k_plain.log 2>&1 && tail -1 gpurun_out/r2k_plain.log
ncu --set full --import-source on --clock-control none -k regex:odl_sweep_kernel -s 2 -c 1 -f -o gpurun_out/r2k_sweep python tools/profile_auto.py auto 8 > gpurun_out/r2k_ncu_sweep.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:odl_sweep_bdf_kernel -s 2 -c 1 -f -o gpurun_out/r2k_tail python tools/profile_auto.py auto 8 > gpurun_out/r2k_ncu_tail.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:odl_mcmc_kernel -s 1 -c 1 -f -o gpurun_out/r2k_mcmc python tools/profile_auto.py mcmc > gpurun_out/r2k_ncu_mcmc.log 2>&1
ncu --set full --clock-control none -k regex:odl_mcmc_kernel -s 1 -c 1 -f -o gpurun_out/r2k_stream_iteration python tools/profile_auto.py stream_iteration > gpurun_out/r2k_ncu_s1.log 2>&1
ncu --set full --clock-control none -k regex:odl_mcmc_kernel -s 1 -c 1 -f -o gpurun_out/r2k_stream_chain python tools/profile_auto.py stream_chain > gpurun_out/r2k_ncu_s2.log 2>&1
ODL_WATCHDOG_SPINS=1000 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k_bench_launch_list.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-facade --no-configs --no-cold > gpurun_out/r2k_ncu_bench.log 2>&1
ls -la gpurun_out/r2k*
