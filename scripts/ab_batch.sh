for i in 1 2; do
python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batched', round(j['value'],1), round(j['sweep_ms_in_fit'],4), j['host_phase_ms_per_iteration'], j['clocks']['sm_mhz'])"
IHTB_NO_BATCH=1 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('sequential', round(j['value'],1), round(j['sweep_ms_in_fit'],4), j['host_phase_ms_per_iteration'], j['clocks']['sm_mhz'])"
done
