"""Relation-graph regime (BASELINE config 3): rspmm over the relation graph of an FB15k-237-scale KG, several batch sizes."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.sweep import measure
from ultra_b200.synth import SCALES, synth_kg_torch, relation_graph_of
d = torch.device("cuda:0")
n, r, e = SCALES["fb15k237"]
for alpha in (1.0, 0.0):
    ei, et, R = synth_kg_torch(n, r, e, alpha, 1236, d)
    rei, ret, RR, RN = relation_graph_of(ei, et, n, R)
    for B in (16, 64, 256):
        measure("relation-graph B=%d alpha=%.0f" % (B, alpha), rei, ret, RN, RR, B * 64, 6583.2)
