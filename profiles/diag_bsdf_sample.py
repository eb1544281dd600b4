import os, sys, ctypes
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np
import lumenrenderer_b200 as lr
from lumenrenderer_b200 import api
import __graft_entry__ as entry
gold = np.load('tests/golden/bsdf_reference.npz')
ob = entry.oracle_bindings()
def osample(mat, v):
    out = np.empty((v.shape[0], 8), np.float32); m = np.ascontiguousarray(mat, np.float32); v = np.ascontiguousarray(v, np.float32)
    assert ob.debug_sample_bsdf(None, m.ctypes.data, v.ctypes.data, v.shape[0], out.ctypes.data) == 0; return out
with lr.Renderer(width=8, height=8) as g:
    for i in range(gold["mats"].shape[0]):
        v = gold["sample_in"][i]
        got = g.sample_bsdf(gold["mats"][i], v[:, 0:3], v[:, 3:6], v[:, 6:9], v[:, 9:12]); ref = gold["sample_out"][i]; orc = osample(gold["mats"][i], v)
        dref = np.abs(got[:, 3:6] - ref[:, 3:6]).max(); dorc = np.abs(got[:, 3:6] - orc[:, 3:6]).max()
        same_dir = np.array_equal(got[:, 3:6].view(np.uint32), orc[:, 3:6].view(np.uint32))
        sc_ref = (np.abs(got[:, :7] - ref[:, :7]) / np.maximum(np.abs(ref[:, :7]), 1.0)); sc_orc = (np.abs(got[:, :7] - orc[:, :7]) / np.maximum(np.abs(orc[:, :7]), 1.0))
        okref = (np.abs(got[:, :7]-ref[:, :7]) <= 1e-5*np.maximum(np.abs(ref[:, :7]),1)).all(axis=1).mean()
        print(i, 'rough %.3f' % gold["mats"][i][15], 'dir vs ref %.1e vs oracle %.1e bit-identical %s | scaled vs ref max %.1e ok %.2f | vs oracle max %.1e' % (dref, dorc, same_dir, sc_ref.max(), okref, sc_orc.max()))
