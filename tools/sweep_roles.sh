# device time of the batched kernel per config-5 fixture and role count
for f in two_rectangles square circle_tangent arc_length parc_coincident inconsistent underconstrained perpendicular; do
  for r in 1 2 3 4; do EZPZ_B200_ROLES=$r python tools/time_small.py $f ${1:-65536}; done
done
