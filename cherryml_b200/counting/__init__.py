from ._count_transitions import (clear_device_results, count_co_transitions, count_transitions,  # noqa: F401
                                device_result)
