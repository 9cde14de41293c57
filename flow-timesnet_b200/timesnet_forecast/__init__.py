"""B200-native drop-in for the hot path of the ``timesnet_forecast`` package.

Only the TimesBlock forward path is provided (``models.timesnet``, ``losses``
and the forecast helpers of ``predict``); the CSV / training / CLI layers of the
reference are out of scope.
"""
__all__ = ["models", "losses", "predict"]
