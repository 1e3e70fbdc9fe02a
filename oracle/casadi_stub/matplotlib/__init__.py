"""Empty stand-in so the reference's geometry/sdf modules import (TEST INFRASTRUCTURE)."""
