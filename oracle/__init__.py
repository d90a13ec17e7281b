"""TEST INFRASTRUCTURE — the parity oracle. Import only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs. The product package never imports this."""
