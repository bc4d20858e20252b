"""The end-to-end drivers under the reference's module path (``cherryml/estimation_end_to_end``);
they live in ``cherryml_b200._public_api``.  The EM-optimizer driver of the reference wraps
external programs (Historian / XRATE) and is not part of this package."""
from .._public_api import (  # noqa: F401
    coevolution_end_to_end_with_cherryml_optimizer,
    lg_end_to_end_with_cherryml_optimizer,
)

CHERRYML_TYPE = "cherry++"
