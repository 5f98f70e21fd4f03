for v in default 0 256:2 384:3 384:4 256:4; do
  if [ "$v" = default ]; then unset AVSR_B200_W2SPLIT; else export AVSR_B200_W2SPLIT=$v; fi
  r=$(timeout 100 python bench.py --no-extras --steps 30 --warmup 5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['value']))")
  echo "W2SPLIT=$v: $r"
done
