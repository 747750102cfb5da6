for v in 1 2 1 2; do
  for args in "--alg fourier --dtype f64 --batch 1024" "--alg dehoog --dtype f64 --batch 512"; do
  NL_B200_F64_CTAS=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-peaks --sustain-s 0 --terms 33 --dobs 1 $args 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('CTAS=$v $args: %.1f Mpts/s ms %.3f fwd %.3f bwd %.3f' % (d['value']/1e6, d['ms_per_step'], d['config']['fwd_ms'], d['config']['bwd_ms']))"
  done
done
