import sys, numpy as np
sys.path.insert(0,"tests"); sys.path.insert(0,".")
from oracle import oracle as O
from raypier_optics_b200 import configs, scene as SC, _abi as A
import raypier_optics_b200.core as core
from raypier_optics_b200.engine import get_engine
for G in (False, True):
    cfg = configs.build(core, "zoo", n=12000, gausslets=G)
    sc = SC.Scene(cfg['face_lists'], cfg['wavelengths'])
    want, wc = O.trace_rays(sc, cfg['rays'], cfg['recursion_limit'], cfg['max_length'])
    eng = get_engine(0); eng.set_scene(sc)
    res = eng.trace(np.ascontiguousarray(cfg['rays']), cfg['max_length'], cfg['recursion_limit'])
    got = res.generations(); print("gausslets", G, res.counts, [len(w) for w in want], res.face_counts.tolist()==wc.tolist())
    for gi,(g,w) in enumerate(zip(got,want)):
        if len(g)!=len(w): print("gen",gi,"size differs"); break
        gb = g['base_ray'] if G else g; wb = w['base_ray'] if G else w
        for f in ('wavelength_idx','parent_idx','end_face_idx','ray_ident','ray_type_id'):
            if not np.array_equal(gb[f], wb[f]): print(" gen",gi,"INT field differs",f, int((gb[f]!=wb[f]).sum()))
        # per face (of the parent's hit: face where this ray ENDS)
        for f in ('origin','direction','normal','E_vector','length','accumulated_path','phase','refractive_index','E1_amp','E2_amp'):
            a = gb[f]; b = wb[f]
            if np.iscomplexobj(a): a=np.stack([a.real,a.imag],-1); b=np.stack([b.real,b.imag],-1)
            a=a.reshape(len(gb),-1); b=b.reshape(len(wb),-1)
            fin = np.isfinite(a)&np.isfinite(b)
            d = np.where(fin, np.abs(a-b)/np.maximum(1,np.abs(b)), 0).max(axis=1)
            tol = 1e-10 if f in ('refractive_index','E1_amp','E2_amp') else 1e-9
            bad = d>tol
            if bad.any():
                # which face created these rays? parent's end face
                if gi>0:
                    pb = (want[gi-1]['base_ray'] if G else want[gi-1])
                    pf = pb['end_face_idx'][wb['parent_idx'][bad]]
                else: pf = np.zeros(bad.sum(),int)
                print(" gen",gi,f,"bad",int(bad.sum()),"max %.2e"%d.max(),"created at faces",np.unique(pf).tolist(),"ending at faces",np.unique(wb['end_face_idx'][bad]).tolist())
        if G:
            for f in ('origin','direction','normal','length'):
                a=g['para_rays'][f].reshape(len(g),-1); b=w['para_rays'][f].reshape(len(w),-1)
                fin=np.isfinite(a)&np.isfinite(b)
                d=np.where(fin,np.abs(a-b)/np.maximum(1,np.abs(b)),0).max(axis=1)
                if (d>1e-9).any(): print(" gen",gi,"para",f,"bad",int((d>1e-9).sum()),"max %.2e"%d.max())
    res.free()
