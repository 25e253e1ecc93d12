from eva_b200.metric import valuation_mse  # noqa: F401
