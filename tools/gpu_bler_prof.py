import sys
sys.path.insert(0, ".")
import torch
from ldpc_3gpp_matlab_b200.bler import BlerSimulator
sim = BlerSimulator(8424, 1/3, 1, iterations=8, early_termination=True, batch=4096, seed=1)
sim.run_batch(-0.3); torch.cuda.synchronize()
sim.run_batch(-0.3); torch.cuda.synchronize()
