"""Task registry + `make` (the role of `plb/envs/__init__.py:5-19`, without gym's global registry).

Ids are `<Task>-v<k>`, k = 1..5 selects the variant block of the scene file and the matching target grid."""
from .env import PlasticineEnv
from .gym_shim import TimeLimit

MAX_EPISODE_STEPS = 50
N_VARIANTS = 5
TASKS = ('Move', 'Torus', 'Rope', 'Writer', 'Pinch', 'Rollingpin', 'Chopsticks', 'Table', 'TripleMove', 'Assembly')


def _spec(task, variant):
    return dict(cfg_path=task.lower() + '.yml', version=variant)


ENVS = {f'{task}-v{k}': _spec(task, k) for task in TASKS for k in range(1, N_VARIANTS + 1)}
ENV_NAMES = list(TASKS)


def make(env_name, nn=False, sdf_loss=10, density_loss=10, contact_loss=1, soft_contact_loss=False, **engine_kwargs):
    """Build the env, set the loss weights (same defaults as the reference), wrap it in a 50-step TimeLimit.
    `engine_kwargs`: dtype ('float32' | 'float64'), device, cfg_overrides."""
    try:
        spec = ENVS[env_name]
    except KeyError:
        raise KeyError(f"unknown env id '{env_name}' (known: {', '.join(sorted(ENVS)[:3])}, ...)") from None
    env = PlasticineEnv(nn=nn, **spec, **engine_kwargs)
    env.taichi_env.loss.set_weights(sdf=sdf_loss, density=density_loss, contact=contact_loss, is_soft_contact=soft_contact_loss)
    return TimeLimit(env, MAX_EPISODE_STEPS)
