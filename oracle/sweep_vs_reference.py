"""TEST INFRASTRUCTURE: differential sweep, C oracle vs the live reference env, over every
shipped domain (ui/domains/*.json) -- run in the build container only.
Usage: python oracle/sweep_vs_reference.py [steps] [seed]"""
import glob, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import ref_loader as rl
from oracle.c_oracle import OracleEnv


def compare(xy, seed, T, name):
    try:
        t = rl.TracedEnv(xy)
    except Exception as ex:  # the reference cannot even reset on this domain
        o = OracleEnv(xy)
        return dict(name=name, ok=True, steps=0, crashed='reset: ' + repr(ex), bad=[], elements=0, episodes=0,
                    oracle_flag=o.crashed)
    o = OracleEnv(xy, original_area=float(t.env.original_area))
    acts = rl.action_stream(seed, T)
    bad = []

    def chk(step, what, a, b):
        ok = np.array_equal(a, b) if isinstance(a, np.ndarray) or isinstance(b, np.ndarray) else a == b
        if not ok:
            bad.append((step, what, a, b))
        return ok

    st = t.state()
    chk(-1, 'obs', t.obs, o.obs()); chk(-1, 'ref', st['ref_index'], o.ref_index)
    chk(-1, 'range', tuple(t.env.estimated_area_range), tuple(o.area_range()))
    nsucc = neps = 0
    for i, a in enumerate(acts):
        try:
            r = t.step(a)
        except Exception as ex:  # the reference itself raised (ZeroDivisionError etc.)
            return dict(name=name, ok=not bad, steps=i, crashed=repr(ex), bad=bad, elements=nsucc, episodes=neps,
                        oracle_flag=o.crashed)
        obs, rew, te, tr, info = o.step(a)
        st = r['pre_reset_state']
        chk(i, 'reward', r['reward'], rew); chk(i, 'term', r['terminated'], te); chk(i, 'trunc', r['truncated'], tr)
        chk(i, 'nel', r['n_elements'], o.n_elements)
        if not r['obs_none']:
            bids, bxy = o.boundary()
            chk(i, 'ids', st['ids'], bids.tolist()); chk(i, 'xy', st['xy'], bxy); chk(i, 'ref', st['ref_index'], o.ref_index)
            ids, keys = o.candidates()
            chk(i, 'cand', st['candidates'], list(zip(ids.tolist(), keys.tolist())))
            chk(i, 'area', st['current_area'], o.current_area); chk(i, 'failed', st['failed_num'], o.failed_num)
            chk(i, 'obs', r['terminal_obs'] if (te or tr) else r['obs'], obs)
        else:
            chk(i, 'obsnone', True, obs is None)
        if bad:
            break
        nsucc += r['success']
        if te or tr:
            neps += 1
            o.reset()
            chk(i, 'reset obs', r['obs'], o.obs())
    return dict(name=name, ok=not bad, steps=T, crashed=None, bad=bad, elements=nsucc, episodes=neps,
                oracle_flag=o.crashed)


if __name__ == '__main__':
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    names = sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(rl.REFERENCE_ROOT, 'ui/domains/*.json')))
    nbad = 0
    start = sys.argv[3] if len(sys.argv) > 3 else ''
    for nm in names:
        if nm < start:
            continue
        try:
            xy = rl.load_domain_xy(nm)
        except Exception as ex:
            print('skip', nm, ex); continue
        res = compare(xy, seed, T, nm)
        nbad += not res['ok']
        print(f"{nm:22s} n={len(xy):4d} ok={res['ok']} steps={res['steps']} el={res['elements']} eps={res['episodes']} "
              f"ref_crash={res['crashed']} oracle_flag={res['oracle_flag']}", flush=True)
        for b in res['bad'][:3]:
            print('   MISMATCH', b[0], b[1], b[2], b[3])
    print('domains with mismatches:', nbad)
