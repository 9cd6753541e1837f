import sys
sys.path[:0]=['coeb-slam_b200/python']
import numpy as np, coeb_b200 as cb
from coeb_b200 import synth
ex=cb.Extractor()
k,d=ex.extract(synth.make_frame(0))
print(len(k))
