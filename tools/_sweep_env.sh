for g in 0 1 2; do
echo "== geom $g (0 = automatic)"
SEQWIN_AGG_EDGE_GEOM=$g python tools/sweep_kw.py --genomes 500 --reps 4 --kw 21:200,21:50,21:10,63:20 2> gpurun_out/tmp.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); s=d.get('stage_ms',{}); print(d['k'],d['w'],d.get('error'),{k:round(v,2) for k,v in s.items() if k in ('total_ms','sort_nodes_ms','nodes_ms','edges_ms')})
"
done
