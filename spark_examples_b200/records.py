"""Record types entering the hot path -- mirrors of the reference's case classes
(src/main/scala/com/google/cloud/genomics/spark/examples/rdd/VariantsRDD.scala:46-54 and
VariantsPca.scala:288)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence


@dataclass(frozen=True)
class Call:
    """rdd/VariantsRDD.scala:46-48.  genotype: allele indices, 0 = reference, > 0 = alternate, -1 = no-call."""
    callsetId: str
    callsetName: str = ""
    genotype: Sequence[int] = ()
    genotypeLikelihood: Optional[Sequence[float]] = None
    phaseset: str = ""
    info: Dict[str, List[str]] = field(default_factory=dict)


@dataclass(frozen=True)
class Variant:
    """rdd/VariantsRDD.scala:51-54."""
    contig: str
    id: str = ""
    names: Optional[List[str]] = None
    start: int = 0
    end: int = 0
    referenceBases: str = ""
    alternateBases: Optional[List[str]] = None
    info: Dict[str, List[str]] = field(default_factory=dict)
    created: int = 0
    variantSetId: str = ""
    calls: Optional[Sequence[Call]] = None


@dataclass(frozen=True)
class CallData:
    """VariantsPca.scala:288."""
    hasVariation: bool
    callsetId: int
