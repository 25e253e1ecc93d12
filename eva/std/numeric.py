from eva_b200.std.numeric import horizontal_sum  # noqa: F401
