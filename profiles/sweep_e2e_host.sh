for cfg in "131072 2" "131072 3" "131072 4" "131072 6"; do
  set -- $cfg
  KIN_HOST_CHUNK=$1 KIN_HOST_FILL_THREADS=$2 python bench.py --skip-callers --skip-cpu --skip-variants --skip-north-star > gpurun_out/tmp_e2e.json 2>/dev/null
  python -c "
import json
d=json.loads(open('gpurun_out/tmp_e2e.json').read().strip().splitlines()[-1])
print('chunk $1 threads $2: e2e %.3e all-rows %.3e pcie_frac %.3f' % (d['e2e']['value'], d['e2e']['value_all_rows_over_pcie'], d['e2e']['pcie_frac']))"
done
