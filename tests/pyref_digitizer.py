"""A second, independent restatement of the reference's digitizer chain for SMALL cases: plain Python loops that follow the
host sequence of gPET.cu:385-424 and the kernels of gPET_kernals.cu:607-698 statement by statement -- in particular the
dead-time kernel is walked literally (chain starts, fp32 `tdead`, the walker running on through survivors), where the C
oracle (oracle/gpet_oracle.c) uses the closed form of SURVEY 8a D7.  TEST INFRASTRUCTURE: it checks the oracle
(tests/test_oracle_and_host.py), nothing else imports it.

What is idealised: every chain start is decided on the list as it is before any kill, and every walker works on its own
copy of that list (on the GPU all walkers run at once; SURVEY D7 shows that for times below ~4e6 us the kill flags do not
depend on how they interleave).  Ties of the unstable std::sort are broken by the position in the input list, as in the
oracle.  Blur is left out (R = 0 leaves E unchanged)."""
import numpy as np

MAXT = 1e20
f32 = np.float32


def _sort_t(e, orig, n):
    """quicksort_h(events, 0, n, 3) (detector.cu:354-367): the first n records by t."""
    k = np.lexsort((orig[:n], e["t"][:n]))
    e[:n] = e[:n][k]
    orig[:n] = orig[:n][k]


def _energywindow(e, n, lo, hi):
    """energywindow (gPET_kernals.cu:641-656): returns the number of events it removed."""
    num = 0
    for i in range(n):
        if e["E"][i] < f32(lo) or e["E"][i] > f32(hi):
            e["t"][i] = MAXT
            num += 1
    return num


def _deadtime(e, n, interval, deadtype):
    """deadtime (gPET_kernals.cu:657-698) on a list ordered by (siten, t); returns the kill flags."""
    site = e["siten"][:n].copy()
    t0 = e["t"][:n].copy()
    iv = float(f32(interval))
    killed = np.zeros(n, bool)
    for start in range(n):
        if not (start == 0 or site[start] != site[start - 1] or t0[start] > t0[start - 1] + iv):
            continue
        t = t0.copy()                       # this walker's view
        current = start
        i = current + 1
        tdead = f32(t[start])
        while i < n:
            while site[i] == site[current] and t[i] < float(f32(tdead + f32(interval))):
                if not deadtype:
                    tdead = f32(t[i])       # paralyzable
                t[i] = MAXT
                killed[i] = True
                i += 1
                if i == n:
                    break
            if i == n:
                break
            if site[i] != site[i - 1] or t[i] > t[i - 1] + iv:
                break
            current = i
            tdead = f32(t[current])
            i += 1
    return killed


def digitize(ev, d):
    """adder.dat-like list -> (singles, counts[4], list of (a, b) index pairs into singles).  d: the keys of
    tests/parity.make_digi_params."""
    e = ev.copy()
    n = e.size
    orig = np.arange(n)
    counts = [n]
    # gPET.cu:393-397
    cnt = n - _energywindow(e, n, d["threshold_eV"], 2000000)
    _sort_t(e, orig, n)
    counts.append(cnt)
    # gPET.cu:400-406
    if d["dead_level"] != 3:
        for i in range(cnt):                # setSitenum, gPET_kernals.cu:607-640
            if d["dead_level"] == 0:
                e["siten"][i] = 0
            elif d["dead_level"] == 1:
                e["siten"][i] = e["pann"][i]
            elif d["dead_level"] == 2:
                e["siten"][i] = e["pann"][i] * d["moduleN"] + e["modn"][i]
    k = np.lexsort((orig[:cnt], e["t"][:cnt], e["siten"][:cnt]))    # orderevents, detector.cu:369-385
    e[:cnt] = e[:cnt][k]
    orig[:cnt] = orig[:cnt][k]
    killed = _deadtime(e, cnt, d["dead_time_us"], d["dead_type"])
    e["t"][:cnt][killed] = MAXT
    _sort_t(e, orig, cnt)
    cnt -= int(killed.sum())
    counts.append(cnt)
    # gPET.cu:418-423
    removed = _energywindow(e, cnt, d["ewin_min"], d["ewin_max"])
    _sort_t(e, orig, cnt)
    cnt -= removed
    counts.append(cnt)
    s = e[:cnt].copy()
    # coincidence sorter: a window of W is opened by the first single that is not inside an earlier
    # window; policy 0 keeps the windows that hold exactly two singles, policy 1 pairs the opener with every single of
    # its window; a pair needs a cyclic panel distance of at least coinc_min_panel_diff
    pairs = []
    W = float(f32(d["coinc_window_us"]))
    if W > 0:
        def ok(a, b):
            if d["coinc_min_panel_diff"] <= 0:
                return True
            dist = abs(int(s["pann"][a]) - int(s["pann"][b]))
            if d["npanels"] > 0:
                dist = min(dist, d["npanels"] - dist)
            return dist >= d["coinc_min_panel_diff"]
        a = 0
        while a < cnt:
            inside = [b for b in range(a + 1, cnt) if s["t"][b] < s["t"][a] + W]
            good = [b for b in inside if ok(a, b)]
            if d["coinc_policy"] == 0:
                if len(inside) == 1 and len(good) == 1:
                    pairs.append((a, good[0]))
            else:
                pairs += [(a, b) for b in good]
            a += len(inside) + 1
    return s, counts, pairs
