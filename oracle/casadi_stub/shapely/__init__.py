"""Empty stand-in so the reference's core/sdf/casadi.py imports (TEST INFRASTRUCTURE)."""
