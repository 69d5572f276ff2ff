import sys; sys.path.insert(0,'/root/repo')
import numpy as np
import falling_sand_engine_b200 as fse
from falling_sand_engine_b200 import worldgen as G, materials as M
table=M.default_materials()
ctx=fse.Context(0,table)
W,H=1536,1024
cells=G.sparse_band(table,W,H,0,H,seed=11,pockets=6)
ga=fse.World(ctx,W,H); ga.write_rect(0,0,cells); ga.active_enable(True)
import ctypes as C
for t in range(60):
    ga.tick(t); ga.particles_tick()
    if t%4==2: ga.tick_temperature()
    if t%6==5:
        a,tot=ga.active_stats()
        h=(C.c_uint8*(tot))()
        print(t,a,tot, ga.particles_count(), int(ga.read_all()['moved'].sum()))
c=ga.read_all()
mats=np.unique(c['mat'][128:-128,128:-128],return_counts=True); print(mats)
