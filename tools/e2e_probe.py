import ctypes as C, time, torch, sys
sys.path.insert(0, ".")
import isaacgymenv_b200
from isaacgymenv_b200 import _lib
lib = _lib.load()
env = isaacgymenv_b200.make(seed=1, task="Anymal", num_envs=4096, sim_device="cuda:0", rl_device="cuda:0", headless=True)
n = 4096
h_act = (2 * torch.rand(n, 12) - 1).pin_memory()
d_act = h_act.cuda()
h_obs = torch.empty(n, 48).pin_memory(); h_rew = torch.empty(n).pin_memory(); h_rs = torch.empty(n, dtype=torch.int64).pin_memory(); h_to = torch.empty(n, dtype=torch.int64).pin_memory()
h_all = torch.empty(n * 48 + n + 2 * n * 2).pin_memory()
d_all = torch.empty(n * 48 + n + 2 * n * 2, device="cuda")
st = torch.cuda.current_stream(); sp = C.c_void_p(st.cuda_stream)
def step(a): _lib.check(lib.b2g_task_step(env.sim.handle, C.c_void_p(a.data_ptr()), sp))
def t(fn, k=300):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e6
def a(): step(d_act); st.synchronize()
def b(): d_act.copy_(h_act, non_blocking=True); step(d_act); st.synchronize()
def c(): d_act.copy_(h_act, non_blocking=True); step(d_act); h_obs.copy_(env.obs_clamped, non_blocking=True); st.synchronize()
def d(): d_act.copy_(h_act, non_blocking=True); step(d_act); h_obs.copy_(env.obs_clamped, non_blocking=True); h_rew.copy_(env.rew_buf, non_blocking=True); h_rs.copy_(env.reset_buf, non_blocking=True); h_to.copy_(env.timeout_buf, non_blocking=True); st.synchronize()
def e(): d_act.copy_(h_act, non_blocking=True); step(d_act); h_all.copy_(d_all, non_blocking=True); st.synchronize()
def f(): _lib.check(lib.b2g_task_step_host(env.sim.handle, C.c_void_p(h_act.data_ptr()), C.c_void_p(h_obs.data_ptr()), C.c_void_p(h_rew.data_ptr()), C.c_void_p(h_rs.data_ptr()), C.c_void_p(h_to.data_ptr()), sp))
def g(): step(h_act); st.synchronize()     # zero-copy actions: kernel reads pinned host memory
def h(): step(h_act); h_all.copy_(d_all, non_blocking=True); st.synchronize()
for name, fn in [("step+sync", a), ("h2d+step+sync", b), ("h2d+step+d2h_obs", c), ("h2d+step+4xd2h", d), ("h2d+step+1 packed d2h", e), ("b2g_task_step_host", f), ("zero-copy actions step+sync", g), ("zero-copy actions + packed d2h", h)]:
    print(f"{name:34s} {t(fn):7.1f} us")
