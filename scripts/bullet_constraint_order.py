"""
Why the constraint rows are ordered "motors, then limits" (DESIGN.md section 4).

Bullet's btMultiBodyDynamicsWorld sorts its multibody constraints by island id with btAlignedObjectArray::quickSort
(Hoare partition, middle pivot, NOT stable) before the solver walks them.  PyBullet adds, per body, one
btMultiBodyJointLimitConstraint per limited joint at load time and then one btMultiBodyJointMotor per joint
(createJointMotors).  Simulating that sort on [button limit, button motor, 12 Kuka limits, 12 Kuka motors] with the
button's island id below the Kuka's (it is loaded first) yields: button motor, button limit, Kuka motors, Kuka limits.
(Restated from memory of Bullet 2.87; the library is not available here.)
"""


def quicksort(a, key):
    def less(x, y):
        return key(x) < key(y)

    def qs(lo, hi):
        i, j = lo, hi
        x = a[(lo + hi) // 2]
        while True:
            while less(a[i], x):
                i += 1
            while less(x, a[j]):
                j -= 1
            if i <= j:
                a[i], a[j] = a[j], a[i]
                i += 1
                j -= 1
            if not (i <= j):
                break
        if lo < j:
            qs(lo, j)
        if i < hi:
            qs(i, hi)
    if len(a) > 1:
        qs(0, len(a) - 1)
    return a


if __name__ == "__main__":
    items = [("button", "L", 0), ("button", "M", 0)] + [("kuka", "L", i) for i in range(12)] + [("kuka", "M", i) for i in range(12)]
    out = quicksort(list(items), lambda t: {"button": 0, "kuka": 1}[t[0]])
    print(" ".join("%s%s%d" % (b[0], k, i) for b, k, i in out))
