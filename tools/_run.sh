(time timeout 120 oracle/_ref/dropin_demo -a 0.3 -b 5 -N 200 -C 50 -I 10 -H 10 -s 7) 2>&1 | tail -25
echo ======
(time timeout 120 oracle/_ref/dropin_check -a 0.3 -b 5 -N 200 -C 20 -I 5 -H 5 -s 7 -STI) 2>&1 | tail -25
