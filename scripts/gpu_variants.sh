#!/bin/bash
# usage (under gpurun): scripts/gpu_variants.sh tagprefix "cfg cfg ..." variant variant ...
pre=$1; cfgs=$2; shift 2
for v in "$@"; do
  echo "== $v"
  HC_LIB=$PWD/hydracore_b200/libhc_$v.so timeout 300 python scripts/gpu_k2_sweep.py ${pre}_$v $cfgs 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception:
        print(ln.rstrip()[-300:]); continue
    print('%-8s prim %5.0f shad %5.0f inc %5.0f incs %5.0f any %5.0f  diffs %s' % (d['cfg'], d['primary_mrays'], d['shadow_mrays'], d['incoherent_mrays'], d['incoherent_shuffled_mrays'], d['incoherent_anyhit_mrays'], [d.get(k) for k in ('diff_primary','diff_shadow','diff_inc','diff_incs','diff_any')]))
"
done
