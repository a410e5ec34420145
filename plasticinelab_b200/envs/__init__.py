"""`plb.envs` surface: the 10 tasks x 5 variants and `make` (`plb/envs/__init__.py:5-19`)."""
from .env import PlasticineEnv
from .gym_shim import TimeLimit

ENV_NAMES = ['Move', 'Torus', 'Rope', 'Writer', 'Pinch', 'Rollingpin', 'Chopsticks', 'Table', 'TripleMove', 'Assembly']
ENVS = {f'{name}-v{i + 1}': dict(cfg_path=f'{name.lower()}.yml', version=i + 1) for name in ENV_NAMES for i in range(5)}
MAX_EPISODE_STEPS = 50


def make(env_name, nn=False, sdf_loss=10, density_loss=10, contact_loss=1, soft_contact_loss=False, **kwargs):
    if env_name not in ENVS:
        raise KeyError(f"unknown env id '{env_name}'")
    env = PlasticineEnv(nn=nn, **ENVS[env_name], **kwargs)
    env.taichi_env.loss.set_weights(sdf=sdf_loss, density=density_loss, contact=contact_loss,
                                    is_soft_contact=soft_contact_loss)
    return TimeLimit(env, MAX_EPISODE_STEPS)
