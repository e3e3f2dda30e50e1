"""Plant models shipped as benchmark fixtures (``qpmpc/systems/``)."""

from .wheeled_inverted_pendulum import WheeledInvertedPendulum

__all__ = ["WheeledInvertedPendulum"]
