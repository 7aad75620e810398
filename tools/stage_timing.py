import sys, time, json
sys.path.insert(0,'.')
import numpy as np
from final184_b200 import api as A, scene as S
from final184_b200.fixture import frame_inputs
N,W,H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sc = S.get_scene()
main, shadow, voxel = (S.fixture_constants(n) for n in ('main','shadow','voxel'))
t=time.time(); fi = frame_inputs(sc, main, shadow, W, H, 2048); print('fixture s', time.time()-t, flush=True)
res = {}
for mode in (A.MODE_NORTHSTAR, A.MODE_REFERENCE):
    g = A.VoxelGI(N, W, H, mode)
    g.upload_scene(sc)
    for s,k in ((A.SLOT_DEPTH,'depth'),(A.SLOT_NORMALS,'normals'),(A.SLOT_SHADOW,'shadow'),(A.SLOT_MATERIAL,'material')): g.upload(s, fi[k])
    k = A.trace_constants_c(main, shadow, voxel, W, H, 0, True)
    for it in range(4):
        g.voxelize(voxel)
        if mode == A.MODE_NORTHSTAR:
            g.inject(k); g.build_mips()
        g.trace_indirect(k); g.gtao(main); g.blur_indirect(k)
        g.sync()
    r = {A.STAGE_NAMES[s]: round(g.stage_ms(s),4) for s in range(A.STAGE_COUNT)}
    r['frags']=g.counter(A.COUNTER_FRAGMENTS); r['samples']=g.counter(A.COUNTER_MARCH_STEPS); r['occ']=g.counter(A.COUNTER_OCCUPIED); r['bricks']=g.counter(A.COUNTER_BRICKS)
    print('mode', mode, json.dumps(r), flush=True)
    g.close()
