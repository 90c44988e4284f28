import os, sys
sys.path.insert(0, os.getcwd())
os.environ["SV_PACK_DEBUG"] = "1"
from splitvae_b200.engine import Engine
e = Engine(model="lgvae", height=64, width=64, batch=256, beta=120.0)
