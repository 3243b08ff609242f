"""debug helper: per-level summary of a STMQR_B200_TRACE file"""
import sys
import pandas as pd
d = pd.read_csv(sys.argv[1])
names = ["build_S", "setup", "assemble", "panel", "update", "finish", "pack", "hpinv"]
d['cls'] = d['class'].map(lambda c: names[c])
pv = d.pivot_table(index='level', columns='cls', values='us', aggfunc='sum').fillna(0)
pv['npanel'] = d[d.cls == 'panel'].groupby('level').size()
pv['panel_avg'] = pv['panel'] / pv['npanel']
pv['update_avg'] = pv['update'] / (pv['npanel'] - 1).clip(lower=1)
print(pv.round(0).to_string())
print(pv.sum().round(0).to_string())
