python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import sys
sys.path.insert(0,'tests')
import workloads as wl
open('/tmp/massive500.md','w').write(wl.massive_problem_text(500, False))
open('/tmp/square.md','w').write(wl.fixture_text('square'))
open('/tmp/tworect.md','w').write(wl.fixture_text('two_rectangles'))
PY
for f in square tworect massive500; do
  echo "== $f (topology cache on)"; ./ezpz_b200/_lib/ezpz-b200 -f /tmp/$f.md | grep -E "Problem size|Iterations|Solved in|solves per"
  echo "== $f (EZPZ_B200_NO_STRUCTURE_CACHE=1: analysis repeated per solve, as the reference does)"; EZPZ_B200_NO_STRUCTURE_CACHE=1 ./ezpz_b200/_lib/ezpz-b200 -f /tmp/$f.md | grep -E "Solved in|solves per"
done
