"""tf.contrib stand-in: only the names the reference imports at module level (unused on the hot path)."""
from . import gan  # noqa: F401


class _Layers(object):
    @staticmethod
    def variance_scaling_initializer(factor=2.0, mode='FAN_IN', uniform=False):
        import tensorflow as tf
        return tf.variance_scaling_initializer(scale=factor, mode=mode.lower(), distribution='uniform' if uniform else 'normal')


layers = _Layers()
