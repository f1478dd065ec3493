for b in attn_prof attn_prof_nomufu; do
  nsys_off=1
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:flash_attn2 -s 1 -c 2 tools/bin/$b 2>&1 | grep -E "gpu__time_duration|done"
done
