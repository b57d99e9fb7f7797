# kin_eval_host on all 8 GPUs of one box at once: host threads per rank (fill, dup)
for cfg in "2 2" "4 4" "3 6" "1 3" "4 8"; do
  set -- $cfg
  KIN_HOST_FILL_THREADS=$1 KIN_HOST_DUP_THREADS=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --steps 5 --warmup 3 --skip-callers --skip-cpu --skip-variants --skip-north-star > gpurun_out/tmp_e2e8.json 2>/dev/null
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/tmp_e2e8.json') if l.startswith('{')][-1])
print('fill $1 dup $2: e2e %.3e all-rows %.3e pcie_frac %.3f check %g' % (d['e2e']['value'], d['e2e']['value_all_rows_over_pcie'], d['e2e']['pcie_frac'], d['e2e']['check_max_abs_diff_vs_device']))"
done
