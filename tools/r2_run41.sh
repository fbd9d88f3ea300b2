cd $GRAFT_REPO_ROOT
for cfg in "80 150" "100 200" "128 200" "160 200" "100 150" "128 150"; do
  set -- $cfg
  echo "== COST=$1 D=$2: $(FIZZ_SPEC_COST=$1 FIZZ_SPEC_D=$2 timeout 120 python tools/timeline_probe.py 96 hybrid 2>&1 | grep '^pass' | tr '\n' ' ')"
done
