"""ORACLE - CPU restatement of the reference's MPC QP path. Test infrastructure only:
nothing under the product package may import from here."""
