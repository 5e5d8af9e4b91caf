"""`python -m spark_examples_b200 [flags]` = VariantsPcaDriver.main (VariantsPca.scala:38-50)."""
from .variants_pca import main

main()
