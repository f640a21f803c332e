import sys, time, ctypes, numpy as np, torch
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parent.parent))
import __graft_entry__ as entry
pkg = entry.load_package(); cap = pkg.capi; api = pkg.default_api()
T, B = 1000, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nx = T * B; n_ids = 3 * nx + (T - 1) * B
is_factor = np.zeros(n_ids, dtype=np.uint8); is_factor[2 * nx:] = 1
ftype = np.zeros(n_ids, dtype=np.int32); ftype[3 * nx:] = 1
xid = np.arange(nx, dtype=np.int64)
ev = np.ascontiguousarray(np.concatenate([np.stack([nx + xid, xid], axis=1).ravel(), np.stack([xid[:nx - B], xid[B:]], axis=1).ravel()]))
ef = np.ascontiguousarray(np.concatenate([np.repeat(2 * nx + xid, 2), np.repeat(3 * nx + xid[:nx - B], 2)]))
store = pkg.SignalStore(api, 2, cap.FAMILY_GAUSS_CANON, cap.F32, 0)
t0 = time.perf_counter()
store.check(api.graph_build(store.h, n_ids, is_factor.ctypes.data_as(cap.u8p), ftype.ctypes.data_as(cap.i32p), len(ev), ev.ctypes.data_as(cap.i64p), ef.ctypes.data_as(cap.i64p)))
one = np.array([1.0])
store.check(api.register_rule(store.h, 0, cap.RULE_GAUSS_OBS, one.ctypes.data_as(cap.f64p), 1))
store.check(api.register_rule(store.h, 1, cap.RULE_GAUSS_RW, one.ctypes.data_as(cap.f64p), 1))
store.check(api.resolve_dependencies(store.h, cap.RESOLVER_DEFAULT_BP))
obs_sig = np.ascontiguousarray(2 * nx + 2 * (2 * xid) + 1, dtype=np.int64)
ev_list = api.prepare_signals(store.h, nx, obs_sig.ctypes.data_as(cap.i64p))
req = api.prepare_request(store.h, nx, xid.ctypes.data_as(cap.i64p))
print("build+prepare s", time.perf_counter() - t0, flush=True)
dev_in = torch.randn(nx, 2, device="cuda"); torch.cuda.synchronize()
stats = cap.UpdateStats()
for it in range(6):
    t0 = time.perf_counter()
    store.check(api.set_values_prepared(store.h, ev_list, ctypes.c_void_p(dev_in.data_ptr()), 1))
    torch.cuda.synchronize(); t1 = time.perf_counter()
    store.check(api.update_marginals_prepared(store.h, req, ctypes.byref(stats)))
    t2 = time.perf_counter()
    print(it, "set ms", 1e3 * (t1 - t0), "update ms", 1e3 * (t2 - t1), "ran", api.last_schedule(store.h), "launches", stats.kernel_launches, flush=True)
import subprocess
def loop(n):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        store.check(api.set_values_prepared(store.h, ev_list, ctypes.c_void_p(dev_in.data_ptr()), 1))
        store.check(api.update_marginals_prepared(store.h, req, ctypes.byref(stats)))
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0) / n
print("loop 5:", loop(5), "loop 200:", loop(200), flush=True)
p = subprocess.Popen(["nvidia-smi", "--query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"], stdout=subprocess.DEVNULL)
time.sleep(0.5)
print("with nvidia-smi -lms 100: loop 5:", loop(5), "loop 200:", loop(200), flush=True)
p.terminate()
