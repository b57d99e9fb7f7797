for cfg in "131072 4 6" "131072 4 0" "131072 4 3" "131072 4 10" "131072 8 8" "65536 4 6" "262144 4 6"; do
  set -- $cfg
  if [ "$3" = "0" ]; then export KIN_HOST_NO_DUP_COPY=1; else unset KIN_HOST_NO_DUP_COPY; fi
  KIN_HOST_CHUNK=$1 KIN_HOST_FILL_THREADS=$2 KIN_HOST_DUP_THREADS=$3 python bench.py --skip-callers --skip-cpu --skip-variants --skip-north-star > gpurun_out/tmp_e2e.json 2>/dev/null
  python -c "
import json
d=json.loads(open('gpurun_out/tmp_e2e.json').read().strip().splitlines()[-1])
print('chunk $1 fill $2 dup $3: e2e %.3e all-rows %.3e pcie_frac %.3f d2h/step %d check %g' % (d['e2e']['value'], d['e2e']['value_all_rows_over_pcie'], d['e2e']['pcie_frac'], d['e2e']['d2h_bytes_per_step'], d['e2e']['check_max_abs_diff_vs_device']))"
done
nproc
