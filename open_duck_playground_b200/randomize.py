"""Domain randomisation plug-in (reference common/randomize.py:26-146).

The reference returns ``(batched_model, in_axes)`` for Brax's vmap wrapper.  Here the per-env model parameters
live in the step library, so the plug-in writes them in place (``oduck_randomize``) and returns the env plus the
names of the randomised fields for callers that inspect ``in_axes``.
"""
FLOOR_GEOM_ID = 0   # kept from the reference; geom 0 is a visual mesh there, so this draw has no physical effect
TORSO_BODY_ID = 1   # the massless "base" body

RANDOMIZED_FIELDS = ("geom_friction", "body_ipos", "dof_frictionloss", "dof_armature", "body_mass", "qpos0",
                     "actuator_gainprm", "actuator_biasprm")


def domain_randomize(env, rng):
    """``rng``: uint32 [N, 2] key data, one key per env (Brax passes ``jax.random.split(key, num_envs)``)."""
    env.randomize(rng)
    return env, {name: 0 for name in RANDOMIZED_FIELDS}
