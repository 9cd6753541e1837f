for c in 32 64 86 128; do
  export COEB_PIPE_CHUNK=$c
  echo -n "pipechunk $c: "; python bench.py --steps 10 --warmup 3 --no-cpu --no-match 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), d['e2e'].get('wall_frames_per_s'))"
done
