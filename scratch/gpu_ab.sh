for v in A B A B; do cp scratch/libs/lib$v.so abm_b200/libabm_b200.so; echo "lib $v"; timeout 100 python scratch/base_probe.py 200; done
