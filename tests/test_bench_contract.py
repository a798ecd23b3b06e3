"""CPU tier: the reference arm of bench.py (`--impl reference`, the reference's CPU path as the oracle port) runs
without a GPU and prints ONE JSON line with the contract's keys; non-zero ranks of a torchrun launch exit without
work.  (The sm_100a arm needs a B200; its line is committed under profiles/.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    env.pop('CUDA_VISIBLE_DEVICES', None)
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                           '--warmup', '0', '--grid', '2', '4', '4'], capture_output=True, text=True, env=env,
                          timeout=600)


def test_reference_arm_prints_one_contract_line():
    p = run()
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    r = json.loads(lines[0])
    assert r['impl'] == 'reference' and r['unit'] == 'panoramas/s' and r['higher_is_better'] is True
    assert r['value'] > 0 and r['steps'] == 1 and r['vs_baseline'] is None and r['data'] == 'synthetic'
    assert r['metric'].startswith('panoramas/sec') and 'workload' in r['config'] and 'model' not in r['config']
    assert r['e2e'] == {'value': r['value'], 'unit': 'panoramas/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cb = r['cpu_baseline']
    assert cb['kind'] == 'port' and cb['value'] == r['value'] and cb['cores'] >= 1 and 'panorama' in cb['sample']
    assert r['gpu_launches'] == 0


def test_reference_arm_other_ranks_exit_quietly():
    p = run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert p.returncode == 0 and p.stdout.strip() == ''


def test_reference_arm_never_maps_the_product_library():
    """the reference arm imported vln_ver_b200 (and so mapped libver_b200.so) to get a state_dict.  It now
    builds its weights from oracle/ref_init.py and loads synth.py as a plain file."""
    code = (
        "import sys, argparse; sys.argv=['bench.py']; sys.path.insert(0, %r)\n"
        "import bench\n"
        "a = argparse.Namespace(grid=[2, 4, 4], mode='train', batch=8)\n"
        "r = bench.cpu_reference(a, steps=1, warmup=0)\n"
        "assert r['value'] > 0 and 'no optimizer step' in r['sample']\n"
        "assert not any(m == 'vln_ver_b200' or m.startswith('vln_ver_b200.') for m in sys.modules), 'package imported'\n"
        "assert 'libver_b200' not in open('/proc/self/maps').read(), 'library mapped'\n"
        "print('clean')\n" % ROOT)
    p = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and p.stdout.strip().endswith('clean'), p.stderr[-2000:]


def test_reference_init_matches_the_module_state_dict_layout():
    """oracle/ref_init.py (weights of the CPU arm) has the product head's keys and shapes on the path."""
    sys.path.insert(0, ROOT)
    import bench
    from oracle import ref_init
    head = bench.build_model((2, 4, 4))
    sd = ref_init.lift_encode_state_dict((2, 4, 4), bench.NCAM, bench.EMBED)
    ref_init.check_against(head.state_dict(), sd)
    on_path = {k for k in head.state_dict() if not k.startswith(('positional_encoding', 'code_weights'))}
    assert on_path == set(sd)
    # the deterministic part of the initialisation is identical
    k = 'transformer.encoder.layers.0.attentions.0.deformable_attention.sampling_offsets.bias'
    assert (head.state_dict()[k] - sd[k]).abs().max() < 1e-6
