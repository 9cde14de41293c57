from .timesnet import (  # noqa: F401
    DataEmbedding, FFTPeriodSelector, InceptionBlock, InceptionBranch, LowRankTemporalContext, PeriodGrouper,
    PeriodGroupResult, PositionalEmbedding, RMSNorm, TimesBlock, TimesNet,
)
