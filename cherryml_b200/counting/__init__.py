from ._count_transitions import count_co_transitions, count_transitions, device_result  # noqa: F401
