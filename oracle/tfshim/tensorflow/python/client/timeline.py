"""tensorflow.python.client.timeline stand-in (tracing; out of scope, never called)."""
