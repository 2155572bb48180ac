run() {
  python bench.py --steps 60 --warmup 8 --no-extras > /tmp/ab.json 2> /tmp/ab.err
  python - "$1" <<'PY'
import sys, json
try:
    d = json.loads(open('/tmp/ab.json').read())
    print(sys.argv[1], round(d["ms_per_step"], 4), d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
    print(open('/tmp/ab.err').read()[-1500:])
PY
}
for i in 1 2 3; do
  run all
  SIGGAN_SIDE_EXTRA=0 run reduce_only
  SIGGAN_SIDE_REDUCE=0 run none
done
