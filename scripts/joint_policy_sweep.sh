for v in default t32w32 t64w16 t16w8 t16w2; do
  echo "== $v"
  if [ $v = default ]; then unset GC_LIBGYMCOOK; else export GC_LIBGYMCOOK=$PWD/build/variants/libgymcook_$v.so; fi
  python scripts/time_joint.py
  GC_STATUS=1 timeout 300 python scripts/time_cfg5.py 1024 60 partial-divider_tl,partial-divider_tl 2>&1 | grep -v single
done
