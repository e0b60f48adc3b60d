# re-runs the probes whose numbers DESIGN.md quotes and keeps their output (copied to profiles/logs/ afterwards)
O=gpurun_out/evidence; mkdir -p $O
(timeout 600 python scratch/soak_compare.py 96 6000 het; timeout 300 python scratch/soak_compare.py 96 600) > $O/r2_soak_vf_symmetric_vs_onesided.log 2>&1
(timeout 300 python scratch/soak_base.py 200 100 20 1500; timeout 300 python scratch/soak_base.py 600 50 5 3000; timeout 300 python scratch/soak_base.py 64 128 40 800) > $O/r2_soak_base_fused_vs_per_phase.log 2>&1
(for n in 10 25 50; do PROBE_N=$n PROBE_P=3 timeout 300 python scratch/base_n100_probe.py 1 16 148 400; done; PROBE_N=50 PROBE_P=3 timeout 200 python scratch/base_n100_probe.py 1024; PROBE_N=100 PROBE_P=3 timeout 300 python scratch/base_n100_probe.py 1 148 600 1024; PROBE_N=10 PROBE_P=50 timeout 200 python scratch/base_n100_probe.py 1) > $O/r2_base_step_path_crossovers.log 2>&1
timeout 300 python scratch/dense_probe.py > $O/r2_dense_probe_sym_variants.log 2>&1
timeout 200 python scratch/hetero_probe.py > $O/r2_hetero_radii_probe.log 2>&1
timeout 200 python scratch/c2_cluster_probe.py > $O/r2_c2_cluster_vs_grid_barrier.log 2>&1
timeout 300 python scratch/c5_tile_probe.py 8 > $O/r2_c5_tile_probe_8.log 2>&1
for f in $O/*.log; do tail -n 2 $f | cut -c1-200; done
