# quick loop: executor parity tests + the host-fed bench leg
set -x
python -m pytest tests/test_gpu_executor.py -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu --no-modes 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']
print('value',d['value'],'e2e',e['value'],e['fov_per_s'],'h2d_gbs',e['h2d_gbs'],'frac',e['frac_of_copy_ceiling'],'ceiling',e['copy_only']['h2d_ceiling_gbs'])
print('int64 over pcie',e['int64_over_pcie']['value'],e['int64_over_pcie']['frac_of_copy_ceiling'],'u16',e['uint16_masks']['value'])"
