# the reference's experiment files, unchanged but for T, through abm_b200.compat (files copied to scratch/ref_exps, not
# committed); usage: run_all_ref_exps.sh [list-of-basenames-file]
cd /root/repo
for f in $(find scratch/ref_exps -name "*.py" | sort); do
  if [ -n "$1" ] && ! grep -qx "$(basename $f)" $1; then continue; fi
  d=$(mktemp -d); cp scratch/ref_exps/dropin.env $d/
  out=$(cd $d && EXPERIMENT_NAME=dropin timeout 300 python /root/repo/scratch/run_ref_exp.py /root/repo/$f 60 2>&1 | tail -3 | tr '\n' ' ' | cut -c1-260)
  n=$(ls $d/abm/data/simulation_data/dropin/batch_* 2>/dev/null | grep -c "^20")
  echo "$(basename $f): runs=$n :: $out"
  rm -rf $d
done
