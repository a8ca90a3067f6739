"""Recipe: vendor the reference's own implementation of the hot path into oracle/_ref/ -- TEST / BASELINE INFRASTRUCTURE.

    python oracle/make_ref.py            # needs /root/reference (build container); writes oracle/_ref/

The reference is pure Python, so "building" it = placing the handful of modules the path consists of where they can be
imported on the GPU box (where /root/reference does not exist).  oracle/_ref/ is git-ignored (no reference source enters
the history) but travels with the tree, like the built .so.  `__graft_entry__.build()` runs this when /root/reference
is present.  Consumers: bench.py's reference arm / cpu_baseline leg (`kind: "reference"`), oracle/gen_golden.py and the
tests that run the reference's trainer over the drop-in modules.  The product package never imports it.

What is vendored, unmodified:
    crowd_nav/policy/{helpers,graph_model,value_estimator,state_predictor,model_predictive_rl}.py   the path itself
    crowd_nav/policy/{gcn,cadrl,multi_human_rl}.py                        the model-free GCN policy
    crowd_nav/utils/{trainer,memory}.py                                   its training-side caller
    crowd_nav/configs/icra_benchmark/*.py                                 the shipped configurations
    crowd_sim/envs/policy/policy.py, crowd_sim/envs/utils/{action,state,utils}.py   boundary types the planner imports
Two package markers are written EMPTY instead of copied (crowd_sim/__init__.py, crowd_sim/envs/__init__.py: the originals
import gym / the simulator, which are not installed), and ONE patched copy is derived:
    crowd_nav/policy/model_predictive_rl_d.py = model_predictive_rl.py with `values.append(value)` (:250) replaced by
    `values.append(float(value))`, the minimal change that lets action_clip (:242-269) run on current torch / numpy
    (np.array over a list of [1,1] tensors builds a 3-D array and argpartition then raises), so that the
    depth > 1 look-ahead (:271-302) can be executed and pinned.
"""
import os
import shutil
import sys

REF = os.environ.get('RGL_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')

FILES = [
    'crowd_nav/__init__.py',
    'crowd_nav/policy/helpers.py', 'crowd_nav/policy/graph_model.py', 'crowd_nav/policy/value_estimator.py',
    'crowd_nav/policy/state_predictor.py', 'crowd_nav/policy/model_predictive_rl.py',
    'crowd_nav/policy/gcn.py', 'crowd_nav/policy/cadrl.py', 'crowd_nav/policy/multi_human_rl.py',
    'crowd_nav/utils/__init__.py', 'crowd_nav/utils/trainer.py', 'crowd_nav/utils/memory.py',
    'crowd_nav/configs/icra_benchmark/__init__.py', 'crowd_nav/configs/icra_benchmark/config.py',
    'crowd_nav/configs/icra_benchmark/mp_separate.py', 'crowd_nav/configs/icra_benchmark/mp_separate_dp.py',
    'crowd_nav/configs/icra_benchmark/mp_detach.py', 'crowd_nav/configs/icra_benchmark/mp_linear.py',
    'crowd_nav/configs/icra_benchmark/rgl.py',
    'crowd_sim/envs/policy/__init__.py', 'crowd_sim/envs/policy/policy.py',
    'crowd_sim/envs/utils/__init__.py', 'crowd_sim/envs/utils/action.py', 'crowd_sim/envs/utils/state.py',
    'crowd_sim/envs/utils/utils.py',
]
EMPTY = ['crowd_sim/__init__.py', 'crowd_sim/envs/__init__.py', 'crowd_nav/configs/__init__.py']
PATCH_SRC, PATCH_DST = 'crowd_nav/policy/model_predictive_rl.py', 'crowd_nav/policy/model_predictive_rl_d.py'
PATCH_OLD, PATCH_NEW = '            values.append(value)\n', '            values.append(float(value))\n'


def make(ref=REF, out=OUT, quiet=False):
    if not os.path.isdir(ref):
        raise FileNotFoundError('reference checkout not found at %s' % ref)
    if os.path.isdir(out):
        shutil.rmtree(out)
    for rel in FILES:
        dst = os.path.join(out, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        src = os.path.join(ref, rel)
        if os.path.exists(src):
            shutil.copyfile(src, dst)
        elif rel.endswith('__init__.py'):
            open(dst, 'w').close()
        else:
            raise FileNotFoundError(src)
    for rel in EMPTY:
        dst = os.path.join(out, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        open(dst, 'w').close()
    text = open(os.path.join(ref, PATCH_SRC)).read()
    if text.count(PATCH_OLD) != 1:
        raise RuntimeError('patch anchor not found exactly once in %s' % PATCH_SRC)
    with open(os.path.join(out, PATCH_DST), 'w') as f:
        f.write(text.replace(PATCH_OLD, PATCH_NEW))
    if not quiet:
        print('vendored %d reference files (+1 patched copy) into %s' % (len(FILES), out))
    return out


def available():
    return os.path.exists(os.path.join(OUT, 'crowd_nav', 'policy', 'graph_model.py'))


def enable():
    """Put oracle/_ref on sys.path (callers: bench.py reference arm, tests, gen_golden).  Returns True if present."""
    if not available():
        return False
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    return True


if __name__ == '__main__':
    make()
