"""CPU oracle (test infrastructure; see oracle/model_oracle.py header). Not imported by the product."""
