from .transforms import FFT, SenseModel  # noqa: F401
from .algorithms import ConjugateGradient  # noqa: F401
