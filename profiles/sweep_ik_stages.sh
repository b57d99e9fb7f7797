# pose-only / constrained batched IK of bench.py under different stage schedules of kin_ik_solve (KIN_IK_STAGES)
for spec in "0" "6,10" "4,6,10" "5,8,12" "8,12" "3,4,6,9" "10"; do
  KIN_IK_STAGES=$spec python bench.py --skip-cpu --skip-variants --skip-north-star --skip-e2e > gpurun_out/tmp_ik.json 2>/dev/null
  python -c "
import json
d=json.loads(open('gpurun_out/tmp_ik.json').read().strip().splitlines()[-1])['callers']
a=d['batched_ik(2^20 targets)']; b=d['batched_ik_collision_constrained']
print('stages $spec: one solve %.2f ms (%.4f conv), 2 restarts %.2f ms (%.4f); constrained one seed %.2f ms, 2 restarts %.2f ms' % (1e3*a['solve_seconds_one_solve'], a['fraction_f_below_1e-6_one_solve'], 1e3*a['solve_seconds_with_2_restarts'], a['fraction_within_1e-3'], 1e3*b['solve_seconds_one_seed'], 1e3*b['solve_seconds_with_2_restarts']))"
done
