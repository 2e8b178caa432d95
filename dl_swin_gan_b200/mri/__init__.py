from .transforms import FFT, SenseModel  # noqa: F401
