#!/bin/bash
# Round 2, session 2: staggered dispatch of the collision class (options stagger_at /
# stagger_share / stagger_min): parity with it on, same-box A/B against the previous build
# (libneutral_b200.head.so), and a sweep of the position on csp.
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py -q -m gpu -x > $O/pytest_stagger.txt 2>&1; echo "pytest exit $?" >> $O/pytest_stagger.txt
tail -3 $O/pytest_stagger.txt
{
for rep in 1 2; do
  for lib in libneutral_b200.head.so libneutral_b200.so; do
    for d in csp split stream; do
      echo "== rep $rep $lib $d: $(NB200_LIB=$lib timeout 200 python tools/step_breakdown.py $d --repeat 3 2>&1 | tail -1)"
    done
  done
done
for at in 30 40 50 55 60; do
  for sh in 50; do
    echo "== stagger_at=$at share=$sh min=40"
    timeout 200 python tools/step_breakdown.py csp --repeat 3 --opts stagger_at=$at,stagger_share=$sh,stagger_min=40 2>&1 | tail -12
  done
done
echo "== stagger_at=50 share=35 min=40"
timeout 200 python tools/step_breakdown.py csp --repeat 3 --opts stagger_at=50,stagger_share=35,stagger_min=40 2>&1 | tail -12
echo "== stagger_at=40 share=50 min=25"
timeout 200 python tools/step_breakdown.py csp --repeat 3 --opts stagger_at=40,stagger_share=50,stagger_min=25 2>&1 | tail -12
echo "== split stagger_at=40 (must not apply: colliders exceed the first wave)"
timeout 200 python tools/step_breakdown.py split --repeat 3 --opts stagger_at=40,stagger_min=40 2>&1 | tail -1
} > $O/stagger_ab.txt 2>&1
grep -E "^==|^total" $O/stagger_ab.txt
