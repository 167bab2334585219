set -u
for v in "" "--no-overlap"; do for e in "AVSR_X=1" "AVSR_LP_CLUSTER=8"; do
  echo "--- $e $v"; env $e timeout -s KILL 300 python bench.py --steps 8 --warmup 3 --skip-cpu-baseline --skip-roofline $v 2>/dev/null | cut -c1-160
done; done
