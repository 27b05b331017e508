"""tf.contrib.gan stand-in (IS / FID helpers of the reference; out of scope, never called)."""
