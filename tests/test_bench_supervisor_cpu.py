"""bench.py's multi-rank supervisor (world >= 4): a configuration whose workers do not finish inside their wall-clock window
is killed and the next, more conservative one is started with a fresh rendezvous port; the first one that finishes wins."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(fake, secs='14,14,14'):
    env = dict(os.environ, WORLD_SIZE='4', RANK='0', LOCAL_RANK='0', MASTER_ADDR='127.0.0.1', MASTER_PORT='29611',
               DLSG_BENCH_FAKE=fake, DLSG_BENCH_TIER_SECONDS=secs, DLSG_BENCH_TIER_GAP='1')
    env.pop('DLSG_BENCH_WORKER', None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--gpus', '4', '--steps', '2', '--warmup', '3'],
                       env=env, capture_output=True, text=True, timeout=180)
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith('{')]
    return r.returncode, lines, r.stderr


def test_first_tier_wins_when_it_finishes():
    rc, lines, _ = _run('none')
    assert rc == 0 and len(lines) == 1
    assert lines[0]['port'] == '29611' and lines[0]['nvls'] is None and '--graph' not in lines[0]['argv']


def test_stalled_tiers_fall_through_to_the_conservative_one():
    rc, lines, err = _run('0,1')
    assert rc == 0 and len(lines) == 1, err
    assert lines[0]['nvls'] == '0' and lines[0]['argv'][-3:] == ['--graph', '0', '--no-gan'] and lines[0]['port'] == str(29611 + 34)
    assert 'tier 0 timed out' in err and 'tier 1 timed out' in err


def test_all_tiers_stalled_reports_an_error_line():
    rc, lines, _ = _run('0,1,2', secs='5,5,5')
    assert rc == 3 and len(lines) == 1 and lines[0]['value'] is None and 'error' in lines[0]


def test_crashed_tier_falls_through_at_once():
    """A child that exits non-zero (a Python error) must not cost its whole wall-clock window (round 1: 290 s x 8 GPUs)."""
    import time
    t0 = time.time()
    rc, lines, err = _run('x0', secs='120,120,120')
    assert rc == 0 and len(lines) == 1, err
    assert lines[0]['nvls'] == '0' and 'tier 0 exited with 7' in err
    assert time.time() - t0 < 60
