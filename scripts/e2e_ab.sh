echo "== zero-copy results"; GC_E2E_ZEROCOPY=1 python scripts/e2e_breakdown.py 2>&1 | head -4
echo "== D2H copy"; python scripts/e2e_breakdown.py 2>&1 | head -4
