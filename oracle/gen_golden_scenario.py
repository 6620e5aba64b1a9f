"""TEST INFRASTRUCTURE ONLY.  tests/golden/scenario.json: the UNMODIFIED reference writer utils/scenario_gen.py:189-254 on seeded inputs
(python oracle/gen_golden_scenario.py, build container only)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims, gen_golden as GG           # noqa: E402
from tests.common import scenario_inputs                 # noqa: E402


def main():
    ref_shims.install()
    sys.modules.setdefault('configargparse', __import__('argparse'))
    from utils.scenario_gen import prepare_output_dict
    model = GG.quiet(ref_shims.make_ref_model, nfuture=5)
    sg, kw, env = scenario_inputs()
    out = prepare_output_dict(sg, 1, env, 0.5, model, **kw)
    with open(os.path.join(ROOT, 'tests', 'golden', 'scenario.json'), 'w') as f:
        json.dump(out, f)
    print('wrote scenario.json keys', list(out.keys()))


if __name__ == '__main__':
    main()
