"""condition — B200-native mirror of the reference's posterior-covariance guidance interface (hot path only):
``measurements.get_operator`` + the four linear operators, ``condition.ConditionOpenAIDenoiser[V2]`` + mat solvers,
``utils.OrthoTransform``."""
