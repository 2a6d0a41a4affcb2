for g in 1 2 3 4; do for t in 150 300 500; do
echo "== geom $g target $t"
SEQWIN_AGG_EDGE_GEOM=$g SEQWIN_AGG_NODE_TARGET=$t SEQWIN_DEBUG_AGG=1 python tools/sweep_kw.py --genomes 500 --reps 3 --kw 21:200,21:10 2> gpurun_out/tmp.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); s=d.get('stage_ms',{}); print(d['k'],d['w'],d.get('error'),{k:round(v,2) for k,v in s.items() if k in ('total_ms','sort_nodes_ms','nodes_ms','edges_ms')})
"; sort -u gpurun_out/tmp.err | cut -c1-150
done; done
