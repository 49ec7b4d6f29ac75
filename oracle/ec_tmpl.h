/* oracle/ec_tmpl.h -- curve arithmetic template, instantiated for G1 (over Fq) and G2 (over Fq2).
 *
 * TEST INFRASTRUCTURE ONLY.  Restates the `curve_impl!` macro of pairing/src/bn256/ec.rs and
 * pairing/src/wnaf.rs formula-for-formula so that the CPU baseline executes the reference's
 * operation counts:
 *   double            ec.rs:301-358  (dbl-2009-l)
 *   add_assign        ec.rs:360-454  (add-2007-bl, with the P==Q / P==-Q / infinity cases)
 *   add_assign_mixed  ec.rs:456-536  (madd-2007-bl)
 *   batch_normalization ec.rs:251-299, into_affine ec.rs:596-629
 *   affine mul (mul_bits) ec.rs:96-103,179-182
 *   wnaf_table / wnaf_form / wnaf_exp  wnaf.rs:4-71
 *
 * Before including define:  F (field type), FN(x) (field function prefix paste), EC(x) (curve
 * prefix paste).
 */
typedef struct { F x, y, z; } EC(jac);
typedef struct { F x, y; int inf; } EC(aff);

static inline EC(jac) EC(jac_zero)(void) { EC(jac) p = {FN(zero)(), FN(one)(), FN(zero)()}; return p; }
static inline EC(aff) EC(aff_zero)(void) { EC(aff) p = {FN(zero)(), FN(one)(), 1}; return p; }
static inline int EC(jac_is_zero)(const EC(jac) *p) { return FN(is_zero)(&p->z); }
static inline EC(jac) EC(from_aff)(const EC(aff) *a) {
    if (a->inf) return EC(jac_zero)();
    EC(jac) p = {a->x, a->y, FN(one)()};
    return p;
}

static void EC(dbl)(EC(jac) *p) {
    if (EC(jac_is_zero)(p)) return;
    F a = p->x; FN(sqr)(&a);
    F b = p->y; FN(sqr)(&b);
    F c = b; FN(sqr)(&c);
    F d = p->x; FN(add)(&d, &b); FN(sqr)(&d); FN(sub)(&d, &a); FN(sub)(&d, &c); FN(dbl)(&d);
    F e = a; FN(dbl)(&e); FN(add)(&e, &a);
    F f = e; FN(sqr)(&f);
    FN(mul)(&p->z, &p->y); FN(dbl)(&p->z);
    p->x = f; FN(sub)(&p->x, &d); FN(sub)(&p->x, &d);
    p->y = d; FN(sub)(&p->y, &p->x); FN(mul)(&p->y, &e);
    FN(dbl)(&c); FN(dbl)(&c); FN(dbl)(&c);
    FN(sub)(&p->y, &c);
}

static void EC(add)(EC(jac) *p, const EC(jac) *o) {
    if (EC(jac_is_zero)(p)) { *p = *o; return; }
    if (EC(jac_is_zero)(o)) return;
    F z1z1 = p->z; FN(sqr)(&z1z1);
    F z2z2 = o->z; FN(sqr)(&z2z2);
    F u1 = p->x; FN(mul)(&u1, &z2z2);
    F u2 = o->x; FN(mul)(&u2, &z1z1);
    F s1 = p->y; FN(mul)(&s1, &o->z); FN(mul)(&s1, &z2z2);
    F s2 = o->y; FN(mul)(&s2, &p->z); FN(mul)(&s2, &z1z1);
    if (FN(eq)(&u1, &u2) && FN(eq)(&s1, &s2)) { EC(dbl)(p); return; }
    if (FN(eq)(&u1, &u2)) { *p = EC(jac_zero)(); return; }
    F h = u2; FN(sub)(&h, &u1);
    F i = h; FN(dbl)(&i); FN(sqr)(&i);
    F j = h; FN(mul)(&j, &i);
    F r = s2; FN(sub)(&r, &s1); FN(dbl)(&r);
    F v = u1; FN(mul)(&v, &i);
    F z = p->z;
    p->x = r; FN(sqr)(&p->x); FN(sub)(&p->x, &j); FN(sub)(&p->x, &v); FN(sub)(&p->x, &v);
    p->y = v; FN(sub)(&p->y, &p->x); FN(mul)(&p->y, &r);
    FN(mul)(&s1, &j); FN(dbl)(&s1);
    FN(sub)(&p->y, &s1);
    FN(add)(&z, &o->z); FN(sqr)(&z); FN(sub)(&z, &z1z1); FN(sub)(&z, &z2z2); FN(mul)(&z, &h);
    p->z = z;
}

static void EC(madd)(EC(jac) *p, const EC(aff) *o) {
    if (o->inf) return;
    if (EC(jac_is_zero)(p)) { p->x = o->x; p->y = o->y; p->z = FN(one)(); return; }
    F z1z1 = p->z; FN(sqr)(&z1z1);
    F u2 = o->x; FN(mul)(&u2, &z1z1);
    F s2 = o->y; FN(mul)(&s2, &p->z); FN(mul)(&s2, &z1z1);
    if (FN(eq)(&p->x, &u2) && FN(eq)(&p->y, &s2)) { EC(dbl)(p); return; }
    F h = u2; FN(sub)(&h, &p->x);
    F hh = h; FN(sqr)(&hh);
    F i = hh; FN(dbl)(&i); FN(dbl)(&i);
    F j = h; FN(mul)(&j, &i);
    F r = s2; FN(sub)(&r, &p->y); FN(dbl)(&r);
    F v = p->x; FN(mul)(&v, &i);
    p->x = r; FN(sqr)(&p->x); FN(sub)(&p->x, &j); FN(sub)(&p->x, &v); FN(sub)(&p->x, &v);
    FN(mul)(&j, &p->y); FN(dbl)(&j);
    p->y = v; FN(sub)(&p->y, &p->x); FN(mul)(&p->y, &r); FN(sub)(&p->y, &j);
    FN(add)(&p->z, &h); FN(sqr)(&p->z); FN(sub)(&p->z, &z1z1); FN(sub)(&p->z, &hh);
}

static inline void EC(neg)(EC(jac) *p) { if (!EC(jac_is_zero)(p)) FN(neg)(&p->y); }

static EC(aff) EC(to_aff)(const EC(jac) *p) {
    if (EC(jac_is_zero)(p)) return EC(aff_zero)();
    F one = FN(one)();
    EC(aff) a; a.inf = 0;
    if (FN(eq)(&p->z, &one)) { a.x = p->x; a.y = p->y; return a; }
    F zinv; FN(inv)(&zinv, &p->z);
    F zp = zinv; FN(sqr)(&zp);
    a.x = p->x; FN(mul)(&a.x, &zp);
    FN(mul)(&zp, &zinv);
    a.y = p->y; FN(mul)(&a.y, &zp);
    return a;
}

/* ec.rs:251-299; `prod` is caller-provided scratch of n elements */
static void EC(batch_normalize)(EC(jac) *v, size_t n, F *prod) {
    F one = FN(one)();
    F tmp = one;
    size_t k = 0;
    for (size_t i = 0; i < n; i++) {
        if (EC(jac_is_zero)(&v[i]) || FN(eq)(&v[i].z, &one)) continue;
        FN(mul)(&tmp, &v[i].z);
        prod[k++] = tmp;
    }
    if (k == 0) return;
    F t; FN(inv)(&t, &tmp); tmp = t;
    for (size_t i = n; i-- > 0;) {
        if (EC(jac_is_zero)(&v[i]) || FN(eq)(&v[i].z, &one)) continue;
        k--;
        F s = (k == 0) ? one : prod[k - 1];
        F newtmp = tmp; FN(mul)(&newtmp, &v[i].z);
        v[i].z = tmp; FN(mul)(&v[i].z, &s);
        tmp = newtmp;
        /* affine transformation (third pass of the reference, fused: same values) */
        F z = v[i].z; FN(sqr)(&z);
        FN(mul)(&v[i].x, &z);
        FN(mul)(&z, &v[i].z);
        FN(mul)(&v[i].y, &z);
        v[i].z = one;
    }
}

/* affine.mul(): MSB-first over all 256 bits of the repr, ec.rs:96-103 */
static EC(jac) EC(mul_bits)(const EC(aff) *p, const uint64_t k[4]) {
    EC(jac) res = EC(jac_zero)();
    for (int i = 255; i >= 0; i--) {
        EC(dbl)(&res);
        if ((k[i / 64] >> (i % 64)) & 1) EC(madd)(&res, p);
    }
    return res;
}

/* wnaf.rs:4-15 */
static void EC(wnaf_table)(EC(jac) *table, EC(jac) base, int window) {
    EC(jac) dbl = base; EC(dbl)(&dbl);
    for (int i = 0; i < (1 << (window - 1)); i++) { table[i] = base; EC(add)(&base, &dbl); }
}
/* wnaf.rs:49-71 */
static EC(jac) EC(wnaf_exp)(const EC(jac) *table, const int64_t *wnaf, int len) {
    EC(jac) result = EC(jac_zero)();
    int found_one = 0;
    for (int i = len - 1; i >= 0; i--) {
        if (found_one) EC(dbl)(&result);
        int64_t n = wnaf[i];
        if (n != 0) {
            found_one = 1;
            if (n > 0) EC(add)(&result, &table[n / 2]);
            else { EC(jac) t = table[(-n) / 2]; EC(neg)(&t); EC(add)(&result, &t); }
        }
    }
    return result;
}
/* Wnaf::new().base(P, 1).scalar(k): window = recommended_wnaf_for_num_scalars(1) = 4
 * (ec.rs:983-997 / 1399-1413, first threshold is 4 scalars -> window 4 for a single scalar) */
static EC(jac) EC(wnaf_mul)(const EC(aff) *p, const uint64_t k[4]) {
    EC(jac) table[8];
    int64_t digits[260];
    EC(wnaf_table)(table, EC(from_aff)(p), 4);
    int len = wnaf_form(digits, k, 4);
    return EC(wnaf_exp)(table, digits, len);
}
