/*
 * tonal_walk_model.c -- TEST INFRASTRUCTURE: CPU model of the tonal labelling.
 *
 * verb()  the reference walk over the linked list of local maxima, statement for statement
 *         (ref: libtoolame-dab/psycho_1.c:267-340, same as oracle/mp2_oracle.c psy1_tonal)
 * fast()  the mask-driven walk k_label runs on the GPU (odr_audioenc_b200/csrc/mp2_kernels.cu): candidates from a
 *         bit mask, neighbourhood test precomputed on the unmodified spectrum where nothing was wiped yet
 * main()  compares spectrum, type flags, list head and list traversal of both on random spectra built to provoke
 *         ties, dense tonals and tonals closer than `run` (the wiped-predecessor cases).  Prints "bad N".
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#define DBMIN (-200.0)
#define T_TONE 20
#define L_LAST (-1)
#define L_STOP (-100)
static double DBT[1000];
static double add_db(double a, double b){ double f=10.0*(a-b); if(f>990.0) return a; if(f<-990.0) return b; int i=(int)f; if(i>=0) return a+DBT[i]; return b+DBT[-i]; }
static int verb(double*x, short*next, signed char*type){
    int last=L_LAST, first=L_LAST, run, last_but_one=L_LAST, tone=L_LAST;
    for(int i=0;i<512;i++){next[i]=L_STOP;type[i]=0;}
    for (int i = 2; i < 500; i++) if (x[i] > x[i-1] && x[i] >= x[i+1]) { type[i]=T_TONE; next[i]=L_LAST; if(last!=L_LAST) next[last]=i; else first=tone=i; last=i; }
    last=L_LAST; first=tone; tone=L_LAST;
    while (first != L_LAST && first != L_STOP) {
        if (first < 3 || first > 500) run = 0; else if (first < 63) run = 2; else if (first < 127) run = 3; else if (first < 255) run = 6; else run = 12;
        double mx = x[first]-7;
        for (int j=2;j<=run;j++) if (mx < x[first-j] || mx < x[first+j]) { type[first]=0; break; }
        if (type[first]==T_TONE) {
            int help=first; if (tone==L_LAST) tone=first;
            while (next[help]!=L_LAST && (next[help]-first)<=run) help=next[help];
            help=next[help]; next[first]=help;
            if ((first-last)<=run) { if (last_but_one!=L_LAST) next[last_but_one]=first; }
            if (first>1 && first<500) { double tmp=add_db(x[first-1],x[first+1]); x[first]=add_db(x[first],tmp); }
            for (int j=1;j<=run;j++){ x[first-j]=x[first+j]=DBMIN; next[first-j]=next[first+j]=L_STOP; type[first-j]=type[first+j]=0; }
            last_but_one=last; last=first; first=next[first];
        } else { if (last!=L_LAST) next[last]=next[first]; int ll=first; first=next[first]; next[ll]=L_STOP; }
    }
    return tone;
}
static int trun(int i){ if(i<3||i>500) return 0; if(i<63) return 2; if(i<127) return 3; if(i<255) return 6; return 12;}
static int ttest(const double*x,int c,int run){ double mx=x[c]-7; for(int j=2;j<=run;j++) if(mx<x[c-j]||mx<x[c+j]) return 0; return 1;}
static int next_bit(const unsigned*m,int p){ int w=(p+1)>>5; if(w>=16) return L_LAST; unsigned bits=m[w]&(~0u<<((p+1)&31)); while(!bits){ if(++w>=16) return L_LAST; bits=m[w];} return w*32+__builtin_ffs(bits)-1;}
static int fast(double*x, short*next, signed char*type){
    unsigned cand[16]={0},t0[16]={0};
    for(int i=0;i<512;i++){next[i]=L_STOP;type[i]=0; int peak=i>=2&&i<500&&x[i]>x[i-1]&&x[i]>=x[i+1]; if(peak){cand[i>>5]|=1u<<(i&31); if(ttest(x,i,trun(i))) t0[i>>5]|=1u<<(i&31);} }
    int tone=L_LAST,last=L_LAST,last_but_one=L_LAST,mod_end=-1; int c=next_bit(cand,-1);
    while(c!=L_LAST){ int run=trun(c); int tonal; if(c-run>mod_end) tonal=(t0[c>>5]>>(c&31))&1; else tonal=ttest(x,c,run);
        if(!tonal){c=next_bit(cand,c);continue;}
        type[c]=T_TONE; if(tone==L_LAST) tone=c; if(last!=L_LAST) next[last]=c; int beyond=next_bit(cand,c+run); next[c]=beyond;
        if((c-last)<=run){ if(last_but_one!=L_LAST) next[last_but_one]=c; }
        if(c>1&&c<500){ double tmp=add_db(x[c-1],x[c+1]); x[c]=add_db(x[c],tmp);} 
        for(int j=1;j<=run;j++){x[c-j]=x[c+j]=DBMIN; next[c-j]=next[c+j]=L_STOP; type[c-j]=type[c+j]=0;}
        mod_end=c+run; last_but_one=last; last=c; c=beyond; }
    if(last!=L_LAST) next[last]=L_LAST;
    return tone;
}
int main(){ for(int i=0;i<1000;i++) DBT[i]=10*log10(1+pow(10.0,i/-100.0));
  srand(1); int bad=0;
  for(int it=0;it<200000;it++){ double x[512],x2[512]; short n1[512],n2[512]; signed char t1[512],t2[512];
    int mode=it%4; for(int i=0;i<512;i++){ double v=(rand()%2000)/20.0; if(mode==1) v=(rand()%300)/20.0+ (i%7==0?20:0); if(mode==2) v=(rand()%40)/2.0; if(mode==3) v=(rand()%3)*8.0; x[i]=x2[i]=v;}
    int a=verb(x,n1,t1), b=fast(x2,n2,t2);
    int d = a!=b || memcmp(x,x2,sizeof x) || memcmp(t1,t2,sizeof t1);
    /* compare list traversal */
    int la[600],lb[600],na=0,nb=0; for(int k=a;k!=L_LAST&&k!=L_STOP&&na<600;k=n1[k]) la[na++]=k; for(int k=b;k!=L_LAST&&k!=L_STOP&&nb<600;k=n2[k]) lb[nb++]=k;
    if(na!=nb||memcmp(la,lb,na*sizeof(int))) d=1;
    if(d){ if(bad<5){ printf("mismatch it=%d heads %d %d na %d nb %d xdiff %d tdiff %d\n",it,a,b,na,nb,memcmp(x,x2,sizeof x)!=0,memcmp(t1,t2,sizeof t1)!=0); for(int k=0;k<na&&k<12;k++) printf(" %d",la[k]); printf(" | "); for(int k=0;k<nb&&k<12;k++) printf(" %d",lb[k]); printf("\n"); for(int i=0;i<512;i++) if(x[i]!=x2[i]||t1[i]!=t2[i]) {printf("  first diff at %d: x %g %g type %d %d\n",i,x[i],x2[i],t1[i],t2[i]);break;} } bad++; }
  }
  printf("bad %d\n",bad); return 0; }
