import sys, os, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, "tests")
from consolver_b200 import _lib
import abi_helpers as ah
lib = _lib.load()
torch.cuda.init()
gen = torch.cuda.default_generators[0]
torch.manual_seed(1234)
B, A, K = 4, 1, 11
seed, off = gen.initial_seed(), gen.get_offset()
print("seed", seed, "off", off)
ref = torch.empty(B * A, K, device="cuda").exponential_(1)
print("consumed", gen.get_offset() - off, "plan", _lib.philox_plan(B * A * K))
table = torch.full((A, K), 1.0 / K, device="cuda")
sd = {"action_values": torch.zeros(A, K, device="cuda"), "mlp.0.weight": torch.zeros(8, 2, device="cuda")}
nthreads, inc = _lib.philox_plan(B * A * K)
q_out = torch.zeros(B * A, K, device="cuda")
out = ah.policy_sample(sd, table, B, 2, 0, 1, rng=_lib.Rng(seed, off, None, nthreads), q_out=q_out)
print("ref ", ref.flatten()[:8].tolist())
print("mine", q_out.flatten()[:8].tolist())
# what uniform would give ref: u = exp(-ref)
print("u_ref ", torch.exp(-ref.flatten()[:8].double()).tolist())
print("u_mine", torch.exp(-q_out.flatten()[:8].double()).tolist())
# try torch.rand to see raw uniforms for same state
torch.manual_seed(1234)
print("rand", torch.rand(8, device="cuda").tolist())
# check whether ref equals mine at some permutation
r, m = ref.flatten(), q_out.flatten()
match = (r.view(-1, 1) == m.view(1, -1)).nonzero()
print("matches (ref idx, mine idx):", match[:12].tolist())
