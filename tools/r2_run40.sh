cd $GRAFT_REPO_ROOT
for cfg in "48 200" "64 200" "80 200" "64 300" "80 400" "100 300"; do
  set -- $cfg
  echo "== COST=$1 D=$2: $(FIZZ_SPEC_COST=$1 FIZZ_SPEC_D=$2 timeout 120 python tools/timeline_probe.py 96 hybrid 2>&1 | grep '^pass' | tr '\n' ' ')"
done
