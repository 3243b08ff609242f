"""debug helper: per-level summary of a STMQR_B200_TRACE file"""
import sys
import pandas as pd
d = pd.read_csv(sys.argv[1])
names = ["build_S", "setup", "assemble", "panel", "update", "finish", "pack", "hpinv", "vextract", "in_vtc", "in_wt",
         "in_apply", "gram", "tmerge", "out_vtc", "out_wt", "out_apply"]
d['cls'] = d['class'].map(lambda c: names[c])
pv = d.pivot_table(index='level', columns='cls', values='us', aggfunc='sum').fillna(0)
cnt = d.pivot_table(index='level', columns='cls', values='us', aggfunc='count').fillna(0)
pd.set_option('display.width', 250)
print("sum of microseconds per level and kernel class")
print(pv.round(0).astype(int).to_string())
print("launch counts")
print(cnt.astype(int).to_string())
print((pv.sum() / 1e3).round(2).to_string())
