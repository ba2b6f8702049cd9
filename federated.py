"""Drop-in launcher mirroring the reference's ``python federated.py ...`` (src/federated.py)."""
from rlr_b200.federated import main

if __name__ == "__main__":
    main()
