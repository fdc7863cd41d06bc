# usage (under gpurun): bash tools/hoist_stress.sh  -> the timed step alone, right after the 16-process reference arm, and beside 16 busy host processes
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-single-frame --no-parity --profile-steps 0"
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["ms_per_step"],2), "ms/step")'
$B 2>/dev/null | python -c "$P" alone
python bench.py --impl reference --steps 1 --warmup 0 > /dev/null 2>&1
$B 2>/dev/null | python -c "$P" after_reference_arm
$B 2>/dev/null | python -c "$P" again
for i in $(seq 1 16); do python -c "
import time
t=time.time()
while time.time()-t<45: pass" & done
sleep 1
$B 2>/dev/null | python -c "$P" beside_16_busy_processes
wait
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity 2>/dev/null | python -c "$P" full_bench_with_e2e_legs
